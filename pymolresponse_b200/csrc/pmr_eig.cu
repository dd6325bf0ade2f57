// One-sided (Hestenes) Jacobi SVD for the excitation-energy eigenproblems.
//
// The reference diagonalises [[A, B], [-B, -A]] (2N x 2N, non-symmetric) with scipy.linalg.eig
// (solvers.py:758-776) or A alone for TDA (:841-856).  Here the RPA problem is reduced on the host
// side (pymolresponse_b200/solvers.py) to the singular values of G = chol(A-B)^T chol(A+B), whose
// squares are the eigenvalues of (A-B)(A+B): the excitation energies ARE the singular values, to
// high relative accuracy.  This file holds the SVD itself: columns of G are orthogonalised pairwise
// by plane rotations applied from the right, G V = U Sigma, V accumulating the rotations.
//
// Layout: Gt and Vt are (N, N) row-major with row p = column p of G (V), so both vectors of a pair
// are contiguous.  One step = N/2 disjoint pairs from a round-robin tournament, one CTA per pair:
// pass 1 reduces alpha = |g_p|^2, beta = |g_q|^2, gamma = g_p . g_q; pass 2 applies the rotation to
// the two rows of Gt and of Vt (both passes stream 4 N doubles, L2-resident up to N of a few
// thousand).  A sweep is N-1 steps (every pair once); the host reads one convergence scalar
// (max |gamma| / sqrt(alpha beta)) per sweep.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "pmr_common.cuh"

namespace {

using namespace pmr;

constexpr int JT = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = (threadIdx.x < JT / 32) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// Round-robin tournament on Ne = N rounded up to even players; player Ne-1 is fixed.
__device__ __forceinline__ void pair_of(int step, int k, int Ne, int& p, int& q) {
  const int m = Ne - 1;
  if (k == 0) {
    p = step % m;
    q = m;
  } else {
    p = (step + k) % m;
    q = (step - k % m + m) % m;
  }
  if (p > q) { const int t = p; p = q; q = t; }
}

__global__ void __launch_bounds__(JT) jacobi_step_kernel(double* __restrict__ Gt,
                                                         double* __restrict__ Vt, int N,
                                                         long long ld, int step, double tol,
                                                         unsigned long long* __restrict__ off) {
  __shared__ double red[JT / 32];
  const int Ne = (N + 1) & ~1;
  int p, q;
  pair_of(step, blockIdx.x, Ne, p, q);
  if (q >= N) return;                       // the padding player of an odd N sits this step out
  double* gp = Gt + (long long)p * ld;
  double* gq = Gt + (long long)q * ld;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < N; i += JT) {
    const double x = gp[i], y = gq[i];
    a = fma(x, x, a);
    b = fma(y, y, b);
    c = fma(x, y, c);
  }
  const double alpha = block_sum(a, red);
  const double beta = block_sum(b, red);
  const double gamma = block_sum(c, red);
  const double scale = sqrt(alpha) * sqrt(beta);
  if (!(scale > 0.0)) return;
  const double rel = fabs(gamma) / scale;
  if (threadIdx.x == 0) atomicMax(off, (unsigned long long)__double_as_longlong(rel));
  if (!(rel > tol)) return;
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  double* vp = Vt + (long long)p * ld;
  double* vq = Vt + (long long)q * ld;
  for (int i = threadIdx.x; i < N; i += JT) {
    const double x = gp[i], y = gq[i];
    gp[i] = cs * x - sn * y;
    gq[i] = sn * x + cs * y;
    const double u = vp[i], w = vq[i];
    vp[i] = cs * u - sn * w;
    vq[i] = sn * u + cs * w;
  }
}

// sigma[p] = |g_p|, gv[p] = g_p . v_p (sign of the eigenvalue when G is symmetric)
__global__ void __launch_bounds__(JT) jacobi_finish_kernel(const double* __restrict__ Gt,
                                                           const double* __restrict__ Vt, int N,
                                                           long long ld, double* __restrict__ sigma,
                                                           double* __restrict__ gv) {
  __shared__ double red[JT / 32];
  const int p = blockIdx.x;
  const double* g = Gt + (long long)p * ld;
  const double* v = Vt + (long long)p * ld;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += JT) {
    a = fma(g[i], g[i], a);
    b = fma(g[i], v[i], b);
  }
  const double s2 = block_sum(a, red);
  const double d = block_sum(b, red);
  if (threadIdx.x == 0) {
    sigma[p] = sqrt(s2);
    gv[p] = d;
  }
}

}  // namespace

extern "C" int pmr_jacobi_svd(double* Gt, double* Vt, int N, long long ld, int max_sweeps,
                              double tol, double* sigma, double* gv, double* scratch,
                              int* sweeps_done, double* final_off, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0 && ld >= N && max_sweeps >= 1 && tol >= 0.0, "jacobi_svd: bad arguments");
  if (sweeps_done) *sweeps_done = 0;
  if (final_off) *final_off = 0.0;
  if (N == 0) return 0;
  PMR_REQUIRE(Gt && Vt && sigma && gv && scratch, "jacobi_svd: null pointer");
  const int Ne = (N + 1) & ~1;
  unsigned long long* off = reinterpret_cast<unsigned long long*>(scratch);
  double rel = 0.0;
  int sweep = 0;
  if (N > 1) {
    for (; sweep < max_sweeps;) {
      PMR_CHECK_CUDA(cudaMemsetAsync(off, 0, sizeof(unsigned long long), st));
      for (int step = 0; step < Ne - 1; ++step) {
        jacobi_step_kernel<<<Ne / 2, JT, 0, st>>>(Gt, Vt, N, ld, step, tol, off);
        PMR_CHECK_LAUNCH("jacobi_step_kernel");
      }
      unsigned long long bits = 0;
      PMR_CHECK_CUDA(cudaMemcpyAsync(&bits, off, sizeof(bits), cudaMemcpyDeviceToHost, st));
      PMR_CHECK_CUDA(cudaStreamSynchronize(st));
      long long sb = (long long)bits;
      memcpy(&rel, &sb, sizeof(rel));
      ++sweep;
      if (!(rel > tol)) break;               // a whole sweep without a rotation: converged
    }
  }
  jacobi_finish_kernel<<<N, JT, 0, st>>>(Gt, Vt, N, ld, sigma, gv);
  PMR_CHECK_LAUNCH("jacobi_finish_kernel");
  if (sweeps_done) *sweeps_done = sweep;
  if (final_off) *final_off = rel;
  if (N > 1 && rel > tol)
    return set_error(PMR_ERR_CONVERGENCE, "jacobi_svd: not converged after %d sweeps (off = %.3e)", sweep, rel);
  return 0;
}

// A <- tril(A) (strictly upper part zeroed): Cholesky factors as plain GEMM operands.
namespace {
__global__ void zero_upper_kernel(double* __restrict__ A, int N, long long lda) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < N && j > i) A[(long long)i * lda + j] = 0.0;
}
}  // namespace

extern "C" int pmr_zero_upper(double* A, int N, long long lda, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0 && N <= 65535 && lda >= N, "zero_upper: bad arguments");
  if (N == 0) return 0;
  PMR_REQUIRE(A, "zero_upper: null pointer");
  zero_upper_kernel<<<dim3((N + 255) / 256, N), 256, 0, st>>>(A, N, lda);
  PMR_CHECK_LAUNCH("zero_upper_kernel");
  return 0;
}

// Column scalings of a row-major (rows, cols) matrix, 32 columns per CTA (coalesced rows).
namespace {
__global__ void __launch_bounds__(256) scale_columns_kernel(double* __restrict__ A, int rows, int cols,
                                                            long long lda,
                                                            const double* __restrict__ scale) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if (j >= cols) return;
  const double f = scale[j];
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < rows; i += 8 * gridDim.y)
    A[(long long)i * lda + j] *= f;
}

// unit 2-norm columns; sign chosen so that the component of largest magnitude is positive
__global__ void __launch_bounds__(256) normalize_columns_kernel(double* __restrict__ A, int rows,
                                                                int cols, long long lda) {
  __shared__ double s_n2[8][33], s_big[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  double n2 = 0.0, big = 0.0;
  if (j < cols)
    for (int i = g; i < rows; i += 8) {
      const double x = A[(long long)i * lda + j];
      n2 = fma(x, x, n2);
      if (fabs(x) > fabs(big)) big = x;
    }
  s_n2[g][lane] = n2;
  s_big[g][lane] = big;
  __syncthreads();
  if (g == 0) {
    for (int k = 1; k < 8; ++k) {
      n2 += s_n2[k][lane];
      if (fabs(s_big[k][lane]) > fabs(big)) big = s_big[k][lane];
    }
    s_n2[0][lane] = n2 > 0.0 ? (big < 0.0 ? -1.0 : 1.0) / sqrt(n2) : 0.0;
  }
  __syncthreads();
  const double f = s_n2[0][lane];
  if (j < cols)
    for (int i = g; i < rows; i += 8) A[(long long)i * lda + j] *= f;
}
}  // namespace

extern "C" int pmr_scale_columns(double* A, int rows, int cols, long long lda, const double* scale,
                                 void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(rows >= 0 && cols >= 0 && lda >= cols, "scale_columns: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  PMR_REQUIRE(A && scale, "scale_columns: null pointer");
  const int gy = std::max(1, std::min(64, rows / 64));
  scale_columns_kernel<<<dim3((cols + 31) / 32, gy), 256, 0, st>>>(A, rows, cols, lda, scale);
  PMR_CHECK_LAUNCH("scale_columns_kernel");
  return 0;
}

extern "C" int pmr_normalize_columns(double* A, int rows, int cols, long long lda, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(rows >= 0 && cols >= 0 && lda >= cols, "normalize_columns: bad arguments");
  if (rows == 0 || cols == 0) return 0;
  PMR_REQUIRE(A, "normalize_columns: null pointer");
  normalize_columns_kernel<<<(cols + 31) / 32, 256, 0, st>>>(A, rows, cols, lda);
  PMR_CHECK_LAUNCH("normalize_columns_kernel");
  return 0;
}
