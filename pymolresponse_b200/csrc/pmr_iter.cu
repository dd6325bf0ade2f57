// Matrix-free iterative path: block preconditioned conjugate gradients on
//
//     H(w) [p; m] = [(1+s) g; (1-s) g],   H(w) = [[K, -w], [-w, M]],  K = A + B_ref, M = A - B_ref
//
// (the GPU replacement for the fixed-point loop of solvers.py:603-637).  The
// orbital Hessian is never formed: sigma = H d is applied from the MO integral blocks,
//
//     A     = c (ia|jb) - (ij|ab) + delta (e_a - e_i)         (explicit_equations_partial.py:34-38)
//     B_ref = -(c (ia|jb) - (ib|ja))                          (:92-96)
//
// with three DMMA GEMMs per application (pmr_dgemm: U = ovov d, W = ovov^{swap(1,3)} d, V = oovv d,
// d an nov x 2k block holding all right-hand sides and frequencies) and the elementwise kernels of
// this file.  Rows of the MO blocks may be sharded over ranks by the occupied index i: every rank
// produces its rows of sigma and the (small) row blocks are all-gathered once per iteration.
//
// Vectors are (nov, 2k) row-major: columns [0,k) hold the p = X + Y parts, columns [k,2k) the
// m = X - Y parts of the k independent systems (one per operator component and frequency, each
// with its own w).  All CG scalars stay on the device.
#include "pmr_common.cuh"

namespace pmr {
namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Zp[(b, j), :] = Z[(j, b), :]   (rows of width ncols; j < no, b < nv)
__global__ void permute_ov_rows_kernel(const double* __restrict__ Z, int no, int nv, int ncols,
                                       double* __restrict__ Zp) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)no * nv * ncols;
  if (e >= total) return;
  const int c = (int)(e % ncols);
  const long long row = e / ncols;           // destination row b * no + j
  const int j = (int)(row % no), b = (int)(row / no);
  Zp[e] = Z[((long long)j * nv + b) * ncols + c];
}

// sigma rows [row0, row0 + rows) from the three GEMM results (local rows) and d (global rows):
//   q_p = cK U_p - V_p + sK W_p + D d_p - w d_m      q_m = cM U_m - V_m + sM W_m + D d_m - w d_p
__global__ void sigma_combine_kernel(const double* __restrict__ U, const double* __restrict__ V,
                                     const double* __restrict__ W, const double* __restrict__ Z,
                                     const double* __restrict__ ediff, const double* __restrict__ w,
                                     long long row0, long long rows, int k, double cK, double sK,
                                     double cM, double sM, double* __restrict__ q) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * k) return;
  const int c = (int)(e % k);
  const long long r = e / k;
  const long long lp = r * 2 * k + c, lm = lp + k;
  const long long gp = (row0 + r) * 2 * k + c, gm = gp + k;
  const double D = ediff[row0 + r], wc = w[c];
  const double up = U ? U[lp] : 0.0, um = U ? U[lm] : 0.0;
  const double wp = W ? W[lp] : 0.0, wm = W ? W[lm] : 0.0;
  q[lp] = cK * up - V[lp] + sK * wp + D * Z[gp] - wc * Z[gm];
  q[lm] = cM * um - V[lm] + sM * wm + D * Z[gm] - wc * Z[gp];
}

// y = P^-1 r with P = [[D, -w], [-w, D]] per excitation: the uncoupled Hessian
__global__ void precond_kernel(const double* __restrict__ r, const double* __restrict__ ediff,
                               const double* __restrict__ w, long long nov, int k,
                               double* __restrict__ y) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nov * k) return;
  const int c = (int)(e % k);
  const long long row = e / k;
  const long long ip = row * 2 * k + c, im = ip + k;
  const double D = ediff[row], wc = w[c];
  const double inv = 1.0 / (D * D - wc * wc);
  const double rp = r[ip], rm = r[im];
  y[ip] = (D * rp + wc * rm) * inv;
  y[im] = (wc * rp + D * rm) * inv;
}

constexpr int DOT_ROWS = 1024;  // rows per block of the first reduction stage

// partial[blk, c] = sum over the block's rows of X[row, c] Y[row, c] + X[row, k + c] Y[row, k + c]
__global__ void __launch_bounds__(256)
coldot_partial_kernel(const double* __restrict__ X, const double* __restrict__ Y, long long nov, int k,
                      double* __restrict__ partial) {
  __shared__ double red[8];
  const long long r0 = (long long)blockIdx.x * DOT_ROWS;
  const long long r1 = min(r0 + (long long)DOT_ROWS, nov);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < k; ++c) {
    double s = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += 256) {
      const long long ip = r * 2 * k + c;
      s += X[ip] * Y[ip] + X[ip + k] * Y[ip + k];
    }
    s = warp_sum_d(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int q = 0; q < 8; ++q) tot += red[q];
      partial[(long long)blockIdx.x * k + c] = tot;
    }
    __syncthreads();
  }
}

__global__ void coldot_final_kernel(const double* __restrict__ partial, int nblocks, int k,
                                    double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * k + c];
  out[c] = s;
}

// alpha_c = rho_c / dq_c;  z += alpha d;  r -= alpha q   (columns whose dq is 0 are converged: skip)
__global__ void cg_update_zr_kernel(double* __restrict__ z, double* __restrict__ r,
                                    const double* __restrict__ d, const double* __restrict__ q,
                                    const double* __restrict__ rho, const double* __restrict__ dq,
                                    long long nov, int k) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nov * 2 * k) return;
  const int c = (int)(e % (2 * k)) % k;
  const double den = dq[c];
  if (den == 0.0) return;
  const double alpha = rho[c] / den;
  z[e] += alpha * d[e];
  r[e] -= alpha * q[e];
}

// beta_c = rho_new_c / rho_old_c;  d = y + beta d
__global__ void cg_update_d_kernel(double* __restrict__ d, const double* __restrict__ y,
                                   const double* __restrict__ rho_new,
                                   const double* __restrict__ rho_old, long long nov, int k) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nov * 2 * k) return;
  const int c = (int)(e % (2 * k)) % k;
  const double den = rho_old[c];
  const double beta = den == 0.0 ? 0.0 : rho_new[c] / den;
  d[e] = y[e] + beta * d[e];
}

}  // namespace
}  // namespace pmr

using namespace pmr;

extern "C" int pmr_iter_permute_ov_rows(const double* Z, int no, int nv, int ncols, double* Zp,
                                        void* stream) {
  PMR_REQUIRE(no >= 0 && nv >= 0 && ncols >= 0, "iter_permute: bad dims");
  const long long total = (long long)no * nv * ncols;
  if (total == 0) return 0;
  PMR_REQUIRE(Z && Zp, "iter_permute: null pointer");
  permute_ov_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(Z, no, nv,
                                                                                           ncols, Zp);
  PMR_CHECK_LAUNCH("permute_ov_rows_kernel");
  return 0;
}

extern "C" int pmr_iter_sigma_combine(const double* U, const double* V, const double* W,
                                      const double* Z, const double* ediff, const double* w,
                                      long long row0, long long rows, int k, double cK, double sK,
                                      double cM, double sM, double* q, void* stream) {
  PMR_REQUIRE(rows >= 0 && k >= 0 && row0 >= 0, "iter_sigma_combine: bad dims");
  if (rows * k == 0) return 0;
  PMR_REQUIRE(V && Z && ediff && w && q, "iter_sigma_combine: null pointer");
  sigma_combine_kernel<<<(unsigned)((rows * k + 255) / 256), 256, 0, as_stream(stream)>>>(
      U, V, W, Z, ediff, w, row0, rows, k, cK, sK, cM, sM, q);
  PMR_CHECK_LAUNCH("sigma_combine_kernel");
  return 0;
}

extern "C" int pmr_iter_precond(const double* r, const double* ediff, const double* w, long long nov,
                                int k, double* y, void* stream) {
  PMR_REQUIRE(nov >= 0 && k >= 0, "iter_precond: bad dims");
  if (nov * k == 0) return 0;
  PMR_REQUIRE(r && ediff && w && y, "iter_precond: null pointer");
  precond_kernel<<<(unsigned)((nov * k + 255) / 256), 256, 0, as_stream(stream)>>>(r, ediff, w, nov, k, y);
  PMR_CHECK_LAUNCH("precond_kernel");
  return 0;
}

extern "C" long long pmr_iter_coldot_workspace(long long nov, int k) {
  return ((nov + DOT_ROWS - 1) / DOT_ROWS) * (long long)(k > 0 ? k : 1);
}

extern "C" int pmr_iter_coldot(const double* X, const double* Y, long long nov, int k, double* out,
                               double* work, void* stream) {
  PMR_REQUIRE(nov >= 0 && k >= 0, "iter_coldot: bad dims");
  if (k == 0) return 0;
  PMR_REQUIRE(X && Y && out && work, "iter_coldot: null pointer");
  const int nblocks = (int)((nov + DOT_ROWS - 1) / DOT_ROWS);
  if (nblocks > 0) {
    coldot_partial_kernel<<<nblocks, 256, 0, as_stream(stream)>>>(X, Y, nov, k, work);
    PMR_CHECK_LAUNCH("coldot_partial_kernel");
  }
  coldot_final_kernel<<<(k + 127) / 128, 128, 0, as_stream(stream)>>>(work, nblocks, k, out);
  PMR_CHECK_LAUNCH("coldot_final_kernel");
  return 0;
}

extern "C" int pmr_iter_update_zr(double* z, double* r, const double* d, const double* q,
                                  const double* rho, const double* dq, long long nov, int k,
                                  void* stream) {
  PMR_REQUIRE(nov >= 0 && k >= 0, "iter_update_zr: bad dims");
  const long long total = nov * 2 * k;
  if (total == 0) return 0;
  PMR_REQUIRE(z && r && d && q && rho && dq, "iter_update_zr: null pointer");
  cg_update_zr_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(z, r, d, q, rho,
                                                                                       dq, nov, k);
  PMR_CHECK_LAUNCH("cg_update_zr_kernel");
  return 0;
}

extern "C" int pmr_iter_update_d(double* d, const double* y, const double* rho_new,
                                 const double* rho_old, long long nov, int k, void* stream) {
  PMR_REQUIRE(nov >= 0 && k >= 0, "iter_update_d: bad dims");
  const long long total = nov * 2 * k;
  if (total == 0) return 0;
  PMR_REQUIRE(d && y && rho_new && rho_old, "iter_update_d: null pointer");
  cg_update_d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(d, y, rho_new,
                                                                                      rho_old, nov, k);
  PMR_CHECK_LAUNCH("cg_update_d_kernel");
  return 0;
}
