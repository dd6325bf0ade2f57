// Dense factorisations and triangular solves built on the DMMA GEMM (pmr_gemm.cu).
//
// Replaces np.linalg.inv / scipy.linalg.cholesky + inv (solvers.py:499-555) and the
// G^-1 . rhs products (solvers.py:401-466) of the reference.  No explicit inverse of the
// Hessian is ever formed on the product path.
//
// Cholesky (right-looking, NB = 128, batched over independent matrices = frequencies):
//   1. potf2_trtri_kernel: one CTA per matrix factors the 128x128 diagonal block in shared
//      memory and writes inv(L_kk) (one warp per column, register-resident substitution);
//   2. panel:     A21 <- A21 * inv(L_kk)^T            (DMMA GEMM, in place, N = 128 = one tile)
//   3. trailing:  A22 <- A22 - A21 A21^T  (lower)     (DMMA GEMM, PMR_GEMM_LOWER)
// Triangular solves use the stored diagonal-block inverses, so every step is a GEMM.
#include <algorithm>
#include <vector>

#include "pmr_common.cuh"

namespace pmr {
namespace {

constexpr int NB = 128;
constexpr int NBLD = NB + 4;   // +4 doubles: the m8n8k4 fragment loads of the in-kernel DMMA products
                               // (lane (g, t) -> row g, k t) fall on 16 distinct 8-byte banks per half warp
static thread_local bool g_lookahead = true;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Diagonal-block kernel: Cholesky of a (<=128)^2 block and the inverse of its factor, one CTA per
// matrix of the batch, everything in shared memory.  The block is padded with the identity to
// 128x128 so every loop is uniform.  Work is organised in 32-wide block columns:
//   factor : warp 0 factors the 32x32 diagonal block D in REGISTERS (row per lane, shuffles for the
//            column broadcasts, rsqrt instead of sqrt + divide) and inverts it right away (column per
//            lane); the rows below are then X = P inv(D)^T -- a small GEMM over ALL threads with a
//            1x4 register tile, not a 32-step dependent substitution on one thread per row -- and the
//            trailing block gets its rank-32 update from 2x2 register tiles;
//   inverse: in place, right to left, from the stored inv(D) blocks:
//            L[below, j] <- -inv(L[below, below]) L[below, j] inv(D_j), 1x4 register tiles.
// This kernel is the latency floor of every panel chain (single-matrix potrf, distributed factor of
// K): 99 us before this layout (ncu: 7.3 barrier-stall cycles per issue, three warps doing the
// substitution while thirteen waited).
constexpr int SB = 32;
constexpr int TLD = SB + 4;
constexpr int POTF2_THREADS = 512;
// development aid: when set (pmr_debug_potf2_timing), thread 0 of CTA 0 stores clock64() at the phase
// boundaries of the diagonal-block kernel (tools/potf2_phases.py prints the differences)
__device__ long long* g_potf2_clock = nullptr;
#define POTF2_TICK(slot)                                                          \
  do {                                                                            \
    if (tid == 0 && blockIdx.x == 0 && g_potf2_clock) g_potf2_clock[slot] = clock64(); \
  } while (0)
// L [NB][NBLD] + T [NB-SB][TLD] + Dinv [NB/SB][SB][TLD] + rdiag [NB]
constexpr size_t POTF2_SMEM =
    sizeof(double) * (NB * NBLD + (NB - SB) * TLD + (NB / SB) * SB * TLD + NB);

__device__ __forceinline__ void potf2_dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// Shared-memory micro GEMM of the diagonal-block kernel on the FP64 TENSOR pipe:
// C(m x n) (op)= sign * A(m x K) B(K x n), one 8 x 8 output tile per warp and pass, DMMA.8x8x4 with
// the fragments read straight from shared memory (one LDS.64 per operand and DMMA; leading
// dimensions = 4 mod 16 doubles make them conflict free).  On B200 the vector FP64 pipe issues a
// warp DFMA every 4 cycles per sub-partition, the tensor pipe does 256 FMA in 16: scalar and
// register-tiled DFMA versions of these products measured 41 k / 13 k cycles for the largest pair.
//   BT     : B is read transposed, B(q, c) = Bp[c * ldb + q]
//   MODE 0 : C = sign * sum          MODE 1 : C -= sum on the lower triangle of a square C (SYRK)
//   kbeg_col: B(q, c) = 0 for q < c (k loop starts at the column tile); kcap_row: A(r, q) = 0 for q > r
//   (k loop ends with the row tile).  m, n multiples of 8, K of 4.
template <bool BT, int MODE>
__device__ __forceinline__ void smem_mm8(const double* __restrict__ Ap, int lda,
                                         const double* __restrict__ Bp, int ldb, double* __restrict__ Cp,
                                         int ldc, int m, int n, int K, int kbeg_col, int kcap_row,
                                         double sign, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int tn = n >> 3, tm = m >> 3;
  for (int tile = warp; tile < tm * tn; tile += POTF2_THREADS / 32) {
    const int tr = tile / tn, tc = tile - tr * tn;
    if (MODE == 1 && tc > tr) continue;
    const int r0 = tr << 3, c0 = tc << 3;
    const int kb = kbeg_col ? c0 : 0;
    const int ke = kcap_row ? min(K, r0 + 8) : K;
    double acc[2] = {0.0, 0.0};
    const double* ap = Ap + (r0 + g) * lda + t;
    const double* bp = BT ? Bp + (c0 + g) * ldb + t : Bp + t * ldb + c0 + g;
#pragma unroll 4
    for (int q0 = kb; q0 < ke; q0 += 4) {
      const double a = ap[q0];
      const double b = BT ? bp[q0] : bp[q0 * ldb];
      potf2_dmma(acc, a, b);
    }
    double* c = Cp + (r0 + g) * ldc + c0 + 2 * t;
    if (MODE == 0) {
      c[0] = sign * acc[0];
      c[1] = sign * acc[1];
    } else {
      if (c0 + 2 * t <= r0 + g) c[0] -= acc[0];
      if (c0 + 2 * t + 1 <= r0 + g) c[1] -= acc[1];
    }
  }
}

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_trtri_kernel(double* __restrict__ A, long long lda, long long stride_a, int jb,
                   double* __restrict__ dinv, long long stride_dinv, int* __restrict__ info,
                   int k0) {
  extern __shared__ double sm[];
  double* L = sm;                          // [NB][NBLD]
  double* T = sm + NB * NBLD;              // [NB-SB][TLD]
  double* Dinv = T + (NB - SB) * TLD;      // [NB/SB][SB][TLD]: inverses of the diagonal 32x32 factors
  double* rdiag = Dinv + (NB / SB) * SB * TLD;  // [NB] reciprocals of the diagonal of L
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr unsigned FULL = 0xffffffffu;
  A += (long long)blockIdx.x * stride_a;
  dinv += (long long)blockIdx.x * stride_dinv;
  info += blockIdx.x;
  POTF2_TICK(0);

  // load the lower triangle: eight independent global loads in flight per thread (a dependent
  // load -> shared-store loop pays one DRAM round trip per iteration: 32 of them)
#pragma unroll 1
  for (int it = 0; it < NB * NB / (POTF2_THREADS * 8); ++it) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = tid + (it * 8 + u) * POTF2_THREADS;
      const int r = idx >> 7, c = idx & (NB - 1);
      v[u] = (r == c) ? 1.0 : 0.0;
      if (r < jb && c <= r) v[u] = A[(long long)r * lda + c];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = tid + (it * 8 + u) * POTF2_THREADS;
      L[(idx >> 7) * NBLD + (idx & (NB - 1))] = v[u];
    }
  }
  __syncthreads();
  POTF2_TICK(1);

  // ------------------------------------------------------------------ factorisation
  for (int kb = 0; kb < NB / SB; ++kb) {
    const int c0 = kb * SB;
    double* Dk = Dinv + kb * SB * TLD;
    if (warp == 0) {
      double a[SB];
#pragma unroll
      for (int k = 0; k < SB; ++k) a[k] = L[(c0 + lane) * NBLD + c0 + k];
      int bad = 0;
#pragma unroll
      for (int j = 0; j < SB; ++j) {
        const double d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0.0) && bad == 0) bad = c0 + j + 1;
        const double rl = rsqrt(d);
        const double ljj = d * rl;
        a[j] = (lane == j) ? ljj : a[j] * rl;
        if (lane == j) rdiag[c0 + j] = rl;
#pragma unroll
        for (int k = j + 1; k < SB; ++k) {
          const double vk = __shfl_sync(FULL, a[j], k);  // L[k][j]
          a[k] -= a[j] * vk;                              // meaningful for lane >= k only
        }
      }
#pragma unroll
      for (int k = 0; k < SB; ++k) L[(c0 + lane) * NBLD + c0 + k] = (k <= lane) ? a[k] : 0.0;
      if (lane == 0 && bad != 0 && *info == 0) *info = k0 + bad;
      __syncwarp();
      // inv(D): lane owns column `lane` (forward substitution on the unit vector)
      double s[SB];
#pragma unroll
      for (int r = 0; r < SB; ++r) s[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        const double x = s[q] * rdiag[c0 + q];
        s[q] = x;
#pragma unroll
        for (int r = q + 1; r < SB; ++r) s[r] -= L[(c0 + r) * NBLD + c0 + q] * x;
      }
#pragma unroll
      for (int r = 0; r < SB; ++r) Dk[r * TLD + lane] = s[r];   // Dk[r][c] = inv(D)(r, c), 0 above
    }
    __syncthreads();
    POTF2_TICK(2 + 3 * kb);
    const int r0 = c0 + SB;
    const int nrows = NB - r0;
    // X = P inv(D)^T: X[r][c] = sum_{q <= c} P[r][q] Dk[c][q]  (Dk[c][q] = 0 for q > c: full k range)
    smem_mm8<true, 0>(L + r0 * NBLD + c0, NBLD, Dk, TLD, T, TLD, nrows, SB, SB, 0, 0, 1.0, warp, lane);
    __syncthreads();
    POTF2_TICK(3 + 3 * kb);
    // copy X into the factor and update the trailing block: L[r][c] -= X[r] . X[c] (lower part)
    for (int idx = tid; idx < nrows * SB; idx += POTF2_THREADS) {
      const int rr = idx >> 5, c = idx & 31;
      L[(r0 + rr) * NBLD + c0 + c] = T[rr * TLD + c];
    }
    smem_mm8<true, 1>(T, TLD, T, TLD, L + r0 * NBLD + r0, NBLD, nrows, nrows, SB, 0, 0, 1.0, warp, lane);
    __syncthreads();
    POTF2_TICK(4 + 3 * kb);
  }

  for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
    const int r = idx >> 7, c = idx & (NB - 1);
    if (r < jb && c <= r) A[(long long)r * lda + c] = L[r * NBLD + c];
  }
  __syncthreads();
  POTF2_TICK(14);

  // ------------------------------------------------------------------ inverse of L, in place
  for (int idx = tid; idx < (NB / SB) * SB * SB; idx += POTF2_THREADS) {
    const int kb = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
    L[(kb * SB + r) * NBLD + kb * SB + c] = Dinv[kb * SB * TLD + r * TLD + c];
  }
  __syncthreads();
  for (int j = NB / SB - 2; j >= 0; --j) {
    const int c0 = j * SB, r0 = c0 + SB, nb = NB - r0;
    // T = W[below, below] L[below, j]  (W = the part of inv(L) already in place: lower triangular,
    // so row block r needs k <= r only), then W[below, j] = -T inv(D_j)  (Dj(q, c) = 0 for q < c)
    smem_mm8<false, 0>(L + r0 * NBLD + r0, NBLD, L + r0 * NBLD + c0, NBLD, T, TLD, nb, SB, nb, 0, 1, 1.0,
                       warp, lane);
    __syncthreads();
    smem_mm8<false, 0>(T, TLD, Dinv + j * SB * TLD, TLD, L + r0 * NBLD + c0, NBLD, nb, SB, SB, 1, 0, -1.0,
                       warp, lane);
    __syncthreads();
    POTF2_TICK(15 + (NB / SB - 2 - j));
  }
  for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
    const int r = idx >> 7, c = idx & (NB - 1);
    if (r < jb && c <= r) dinv[(long long)r * NB + c] = L[r * NBLD + c];
  }
  __syncthreads();
  POTF2_TICK(18);
}

constexpr int SHIFT_MAX = 64;  // frequencies per launch (shifts travel as kernel parameters)
struct ShiftPack { double a[SHIFT_MAX]; double b[SHIFT_MAX]; };

__global__ void shifted_combine_kernel(const double* __restrict__ M, const double* __restrict__ W,
                                       int N, long long ldm, long long ldw, int nf,
                                       double* __restrict__ S, long long lds, long long stride_s,
                                       const ShiftPack sh) {
  const int row = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col > row || col >= N) return;
  const double m = M[(long long)row * ldm + col];
  const double w = W ? W[(long long)row * ldw + col] : 0.0;
  for (int f = 0; f < nf; ++f) {
    double v = m + sh.b[f] * w;
    if (col == row) v += sh.a[f];
    S[(long long)f * stride_s + (long long)row * lds + col] = v;
  }
}

// A[c][r] = A[r][c] for r > c: 32 x 32 tiles through shared memory (both sides coalesced)
__global__ void symmetrize_lower_kernel(double* __restrict__ A, int N, long long lda) {
  __shared__ double tile[32][33];
  // blockIdx.x enumerates the tiles (tr >= tc) of the lower triangle row by row
  int tr = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((tr + 1) * (tr + 2) / 2 <= (int)blockIdx.x) ++tr;
  while (tr * (tr + 1) / 2 > (int)blockIdx.x) --tr;
  const int tc = (int)blockIdx.x - tr * (tr + 1) / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
  for (int q = ty; q < 32; q += 8) {
    const int r = tr * 32 + q, c = tc * 32 + tx;
    tile[q][tx] = (r < N && c < N) ? A[(long long)r * lda + c] : 0.0;
  }
  __syncthreads();
  for (int q = ty; q < 32; q += 8) {
    const int r = tc * 32 + q, c = tr * 32 + tx;   // destination (row in the column-tile range)
    if (r < N && c < N && c > r) A[(long long)r * lda + c] = tile[tx][q];
  }
}

// Z[kb*NB + r][kb*NB + c] = dinv[kb][r][c] for every diagonal block kb (grid: nblk x NB rows)
__global__ void place_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ Z,
                                         long long ldz, int N) {
  const int kb = blockIdx.x, r = blockIdx.y;
  const int k0 = kb * NB;
  if (k0 + r >= N) return;
  const double* src = dinv + (long long)kb * NB * NB + (long long)r * NB;
  double* dst = Z + (long long)(k0 + r) * ldz + k0;
  for (int c = threadIdx.x; c < NB && k0 + c < N; c += blockDim.x) dst[c] = src[c];
}

__global__ void copy_block_kernel(const double* __restrict__ src, long long lds,
                                  double* __restrict__ dst, long long ldd, int rows, int cols) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows && c < cols) dst[(long long)r * ldd + c] = src[(long long)r * lds + c];
}

// ------------------------------------------------------------------------------------ LU
constexpr int PB = 64;  // LU panel width

// stage 1 of the pivot search in column j over rows [j, N)
__global__ void lu_pivot_partial_kernel(const double* __restrict__ A, long long lda, int N, int j,
                                        double* __restrict__ pval, int* __restrict__ pidx) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = -1.0;
  int bi = j;
  for (int i = j + blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double v = fabs(A[(long long)i * lda + j]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double v = sv[threadIdx.x + s];
      const int i2 = si[threadIdx.x + s];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { pval[blockIdx.x] = sv[0]; pidx[blockIdx.x] = si[0]; }
}

// stage 2: pick the pivot, record it, swap rows j <-> p over all N columns
__global__ void lu_pivot_swap_kernel(double* __restrict__ A, long long lda, int N, int j,
                                     const double* __restrict__ pval, const int* __restrict__ pidx,
                                     int nparts, int* __restrict__ ipiv, int* __restrict__ info) {
  __shared__ int sp;
  if (threadIdx.x == 0) {
    double best = -1.0;
    int bi = j;
    for (int q = 0; q < nparts; ++q)
      if (pval[q] > best || (pval[q] == best && pidx[q] < bi)) { best = pval[q]; bi = pidx[q]; }
    ipiv[j] = bi;
    if (!(best > 0.0) && *info == 0) *info = j + 1;
    sp = bi;
  }
  __syncthreads();
  const int p = sp;
  if (p == j) return;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const double a = A[(long long)j * lda + c];
    const double b = A[(long long)p * lda + c];
    A[(long long)j * lda + c] = b;
    A[(long long)p * lda + c] = a;
  }
}

// scale column j below the diagonal and update the rest of the panel row-wise (warp per row)
__global__ void lu_panel_update_kernel(double* __restrict__ A, long long lda, int N, int j,
                                       int pend) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int i = j + 1 + blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (i >= N) return;
  const double piv = A[(long long)j * lda + j];
  double* row = A + (long long)i * lda;
  const double l = row[j] / piv;
  for (int c = j + 1 + lane; c < pend; c += 32) row[c] -= l * A[(long long)j * lda + c];
  __syncwarp();
  if (lane == 0) row[j] = l;
}

// U12 <- inv(L11) A12 with L11 unit lower (pb x pb) : thread per column
__global__ void lu_trsm_u12_kernel(double* __restrict__ A, long long lda, int k0, int pb, int N) {
  __shared__ double L11[PB][PB + 1];
  for (int idx = threadIdx.x; idx < pb * pb; idx += blockDim.x) {
    const int r = idx / pb, c = idx - r * pb;
    L11[r][c] = A[(long long)(k0 + r) * lda + k0 + c];
  }
  __syncthreads();
  const int c = k0 + pb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double x[PB];
#pragma unroll 1
  for (int r = 0; r < pb; ++r) {
    double s = A[(long long)(k0 + r) * lda + c];
    for (int q = 0; q < r; ++q) s -= L11[r][q] * x[q];
    x[r] = s;
    A[(long long)(k0 + r) * lda + c] = s;
  }
}

__global__ void lu_apply_pivots_kernel(double* __restrict__ B, long long ldb, int nrhs, int N,
                                       const int* __restrict__ ipiv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nrhs) return;
  for (int j = 0; j < N; ++j) {
    const int p = ipiv[j];
    if (p != j) {
      const double a = B[(long long)j * ldb + c];
      B[(long long)j * ldb + c] = B[(long long)p * ldb + c];
      B[(long long)p * ldb + c] = a;
    }
  }
}

// Solve with the pb x pb diagonal block (unit lower, or upper) for every rhs column in place.
__global__ void lu_diag_solve_kernel(const double* __restrict__ LU, long long lda, int k0, int pb,
                                     double* __restrict__ B, long long ldb, int nrhs, int upper) {
  __shared__ double T[PB][PB + 1];
  for (int idx = threadIdx.x; idx < pb * pb; idx += blockDim.x) {
    const int r = idx / pb, c = idx - r * pb;
    T[r][c] = LU[(long long)(k0 + r) * lda + k0 + c];
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nrhs) return;
  double x[PB];
  if (!upper) {
#pragma unroll 1
    for (int r = 0; r < pb; ++r) {
      double s = B[(long long)(k0 + r) * ldb + c];
      for (int q = 0; q < r; ++q) s -= T[r][q] * x[q];
      x[r] = s;
      B[(long long)(k0 + r) * ldb + c] = s;
    }
  } else {
#pragma unroll 1
    for (int r = pb - 1; r >= 0; --r) {
      double s = B[(long long)(k0 + r) * ldb + c];
      for (int q = r + 1; q < pb; ++q) s -= T[r][q] * x[q];
      s /= T[r][r];
      x[r] = s;
      B[(long long)(k0 + r) * ldb + c] = s;
    }
  }
}

}  // namespace
}  // namespace pmr

using namespace pmr;

extern "C" int pmr_debug_potf2_timing(long long* device_buffer /* >= 19 slots, or NULL to stop */) {
  PMR_CHECK_CUDA(cudaMemcpyToSymbol(g_potf2_clock, &device_buffer, sizeof(device_buffer)));
  return 0;
}

extern "C" void pmr_set_lookahead(int on) { g_lookahead = on != 0; }

extern "C" long long pmr_potrf_dinv_doubles(int N) {
  const long long nblk = (N + NB - 1) / NB;
  return nblk * NB * NB;
}

// Cholesky of the leading N x N block of a (rows x N) matrix, rows >= N; the rows below the square
// are carried through every panel solve and trailing update, so on exit they hold Y^T with
// L Y = (those rows)^T: right-hand sides stored as extra ROWS come out forward-substituted for free
// (the forward pass of pmr_potrs would read the whole factor once more).
extern "C" int pmr_potrf_rows(double* A, int N, int rows, long long lda, int batch,
                              long long stride_a, double* dinv, int* info, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0 && batch >= 1 && batch <= 65535, "potrf: bad N/batch");
  PMR_REQUIRE(rows >= N, "potrf: rows < N");
  PMR_REQUIRE(A && dinv && info, "potrf: null pointer");
  PMR_REQUIRE(lda >= N, "potrf: lda < N");
  const long long dstride = pmr_potrf_dinv_doubles(N);
  PMR_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
  if (N == 0) return 0;
  PMR_CHECK_CUDA(cudaMemsetAsync(dinv, 0, sizeof(double) * dstride * batch, st));
  static thread_local bool configured = false;
  const size_t smem = POTF2_SMEM;
  if (!configured) {
    PMR_CHECK_CUDA(cudaFuncSetAttribute(potf2_trtri_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // Two-level blocking + look-ahead.  128-wide panels are factorised and applied inside an outer
  // block of NBO = 512 columns only; the rest of the matrix sees ONE rank-512 update per outer
  // block (C tile read/written once per 512 columns, K = 512 GEMMs).  That update is split in two:
  //   U1: the columns of the NEXT outer block  -> auxiliary high-priority stream `sa`
  //   U2: everything to the right of it        -> caller's stream `st`
  // so the latency-bound panel work of block J+1 (potf2 + panel TRSM + inner updates, all on `sa`)
  // runs underneath the big U2 of block J instead of between two of them.
  constexpr int NBO = 512;
  static thread_local cudaStream_t sa = nullptr;
  if (!sa) {
    int lo = 0, hi = 0;
    PMR_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PMR_CHECK_CUDA(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi));
  }
  std::vector<cudaEvent_t> events;
  auto new_event = [&](cudaStream_t on) -> cudaEvent_t {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventRecord(e, on);
    events.push_back(e);
    return e;
  };
  const bool lookahead = g_lookahead && N > 2 * NBO;
  cudaStream_t sp = lookahead ? sa : st;  // stream of the panel work

  if (lookahead) PMR_CHECK_CUDA(cudaStreamWaitEvent(sa, new_event(st), 0));
  cudaEvent_t u2_done_prev = nullptr;  // U2 of block J-1 finished (on st)

  // C[r0.., c0..] (lower) -= P[r0.., :] P[c0.., :]^T with P = columns [j0, j0+kw) of L, K = kw,
  // starting at panel column offset koff (rank-128 updates use a 128-wide slice of the block)
  auto syrk = [&](int jblk, int j0, int koff, int kw, int r0, int nrows, int c0, int ncols,
                  cudaStream_t s) -> int {
    double* Cp = A + (long long)r0 * lda + c0;
    const double* Pr = A + (long long)r0 * lda + j0 + koff;
    const double* Pc = A + (long long)c0 * lda + j0 + koff;
    pmr_gemm_t t = make_gemm(nrows, ncols, kw, -1.0, Pr, lda, 1, Pc, 1, lda, 1.0, Cp, lda, 1);
    t.a_b1 = stride_a; t.b_b1 = stride_a;
    (void)jblk;
    t.batch1 = batch; t.c_b1 = stride_a;
    t.flags = PMR_GEMM_LOWER;
    return gemm(t, s);
  };

  int jblk = 0;
  for (int j0 = 0; j0 < N; j0 += NBO, ++jblk) {
    const int jend = std::min(N, j0 + NBO);
    // ---- panel work of this outer block (stream sp)
    for (int k0 = j0; k0 < jend; k0 += NB) {
      const int kb = k0 / NB;
      const int jb = std::min(NB, N - k0);
      const int rest = rows - k0 - jb;   // rows below the diagonal block (extra rows included)
      double* Akk = A + (long long)k0 * lda + k0;
      double* dk = dinv + (long long)kb * NB * NB;
      // inside the outer block the updates are LEFT-looking: this 128-wide block column receives the
      // contributions of all earlier panels of the block in ONE GEMM with k = k0 - j0 (128, 256,
      // 384) just before it is factorised, instead of three rank-128 updates spread over the panels
      // (same flops; the short-k launches were the slow 11 % of the trailing-update class)
      if (k0 > j0) PMR_TRY(syrk(jblk, j0, 0, k0 - j0, k0, rows - k0, k0, jb, sp));
      potf2_trtri_kernel<<<batch, POTF2_THREADS, smem, sp>>>(Akk, lda, stride_a, jb, dk, dstride,
                                                             info, k0);
      PMR_CHECK_LAUNCH("potf2_trtri_kernel");
      if (rest > 0) {
        double* A21 = A + (long long)(k0 + jb) * lda + k0;
        pmr_gemm_t g = make_gemm(rest, jb, jb, 1.0, A21, lda, 1, dk, 1, NB, 0.0, A21, lda, 1);
        g.batch1 = batch; g.a_b1 = stride_a; g.b_b1 = dstride; g.c_b1 = stride_a;
        g.flags = PMR_GEMM_C_ALIASES_A;
        PMR_TRY(gemm(g, sp));
      }
    }
    const int rest = N - jend;          // columns still to be updated
    if (rest <= 0) break;
    const int kw = jend - j0;
    if (!lookahead) {
      PMR_TRY(syrk(jblk, j0, 0, kw, jend, rows - jend, jend, rest, st));
      continue;
    }
    const int jnext = std::min(N, jend + NBO);
    const int w1 = jnext - jend;  // width of the next outer block
    cudaEvent_t panel_done = new_event(sa);
    // U1 (sa): columns of the next outer block, after U2 of the previous block (same region)
    if (u2_done_prev) PMR_CHECK_CUDA(cudaStreamWaitEvent(sa, u2_done_prev, 0));
    PMR_TRY(syrk(jblk, j0, 0, kw, jend, rows - jend, jend, w1, sa));
    // U2 (st): everything to the right of it, after the panel of this block
    const int rest2 = N - jnext;
    PMR_CHECK_CUDA(cudaStreamWaitEvent(st, panel_done, 0));
    if (rest2 > 0) PMR_TRY(syrk(jblk, j0, 0, kw, jnext, rows - jnext, jnext, rest2, st));
    u2_done_prev = new_event(st);
  }
  if (lookahead) PMR_CHECK_CUDA(cudaStreamWaitEvent(st, new_event(sa), 0));
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  return 0;
}

extern "C" int pmr_potrf(double* A, int N, long long lda, int batch, long long stride_a,
                         double* dinv, int* info, void* stream) {
  return pmr_potrf_rows(A, N, N, lda, batch, stride_a, dinv, info, stream);
}

// One outer block column of the blocked Cholesky, on its own: factor columns [j0, j0+512) of A
// (rows j0..N-1 must already carry every update from the blocks to the left).  Building block of the
// multi-GPU factorisation (block columns dealt cyclically to ranks, panels broadcast); the caller
// applies the rank-512 update of the other block columns with pmr_dgemm + PMR_GEMM_LOWER.
extern "C" int pmr_potrf_panel(double* A, int N, long long lda, int j0, double* dinv, int* info,
                               void* stream) {
  cudaStream_t st = as_stream(stream);
  constexpr int NBO = 512;
  PMR_REQUIRE(A && dinv && info, "potrf_panel: null pointer");
  PMR_REQUIRE(N >= 0 && j0 >= 0 && j0 < N && j0 % NBO == 0 && lda >= N, "potrf_panel: bad block");
  const long long dstride = pmr_potrf_dinv_doubles(N);
  static thread_local bool configured = false;
  const size_t smem = POTF2_SMEM;
  if (!configured) {
    PMR_CHECK_CUDA(cudaFuncSetAttribute(potf2_trtri_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int jend = std::min(N, j0 + NBO);
  for (int k0 = j0; k0 < jend; k0 += NB) {
    const int kb = k0 / NB;
    const int jb = std::min(NB, N - k0);
    const int rest = N - k0 - jb;
    double* dk = dinv + (long long)kb * NB * NB;
    if (k0 > j0) {
      // left-looking inside the block column: one update with k = k0 - j0 (see pmr_potrf_rows)
      const double* Pk = A + (long long)k0 * lda + j0;
      pmr_gemm_t t = make_gemm(N - k0, jb, k0 - j0, -1.0, Pk, lda, 1, Pk, 1, lda, 1.0,
                               A + (long long)k0 * lda + k0, lda, 1);
      t.flags = PMR_GEMM_LOWER;
      PMR_TRY(gemm(t, st));
    }
    potf2_trtri_kernel<<<1, POTF2_THREADS, smem, st>>>(A + (long long)k0 * lda + k0, lda, 0, jb, dk,
                                                       dstride, info, k0);
    PMR_CHECK_LAUNCH("potf2_trtri_kernel");
    if (rest > 0) {
      double* A21 = A + (long long)(k0 + jb) * lda + k0;
      pmr_gemm_t g = make_gemm(rest, jb, jb, 1.0, A21, lda, 1, dk, 1, NB, 0.0, A21, lda, 1);
      g.flags = PMR_GEMM_C_ALIASES_A;
      PMR_TRY(gemm(g, st));
    }
  }
  return 0;
}

// ---- fused in-chunk substitutions ---------------------------------------------------------------
// Inside one 512-row chunk a triangular solve is four dependent 128-row steps (multiply by the
// inverse diagonal block, update the rows below / above inside the chunk).  As GEMM launches that is
// eight launches of a few microseconds of work each, and with a handful of right-hand-side columns
// the whole solve was launch bound (252 launches, 4.2 ms for 3 columns at N = 7168).  Here one CTA
// takes four columns of one matrix through the whole chunk with the running right-hand side in
// shared memory: warp-per-row dot products with the lanes along k (coalesced rows of L and of the
// inverse blocks) for the forward pass, lanes along the output row for the transposed accesses of
// the backward pass.  The rank-512 update that carries the chunk into the rest of the right-hand
// side stays a DMMA GEMM.
constexpr int TS_NC = 4;
constexpr int TS_THREADS = 512;
constexpr int TS_CHUNK = 512;
static thread_local bool g_fused_chunks = true;

__global__ void __launch_bounds__(TS_THREADS) chunk_forward_kernel(
    const double* __restrict__ L, long long lda, long long stride_l, const double* __restrict__ dinv,
    long long dstride, const double* __restrict__ B, long long ldb, long long stride_b,
    double* __restrict__ W, long long ldw, long long wstride, int nrhs, int k_first, int cend) {
  __shared__ double vec[TS_CHUNK][TS_NC];
  __shared__ double ybuf[NB][TS_NC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = blockIdx.x * TS_NC;
  const int ncol = min(TS_NC, nrhs - col0);
  const long long b = blockIdx.y;
  L += b * stride_l; dinv += b * dstride; B += b * stride_b; W += b * wstride;
  const int rows = cend - k_first;
  for (int idx = tid; idx < rows * TS_NC; idx += TS_THREADS) {
    const int r = idx / TS_NC, c = idx - r * TS_NC;
    vec[r][c] = c < ncol ? B[(long long)(k_first + r) * ldb + col0 + c] : 0.0;
  }
  __syncthreads();
  for (int k0 = k_first; k0 < cend; k0 += NB) {
    const int jb = min(NB, cend - k0);
    const int off = k0 - k_first;
    const double* D = dinv + (long long)(k0 / NB) * NB * NB;
    // y = D * vec[off : off + jb]   (D lower triangular: k <= r).  Four rows per warp at a time:
    // all 16 loads of a lane are issued before the first use (the loop is latency, not flop, bound)
    for (int r4 = warp * 4; r4 < jb; r4 += (TS_THREADS / 32) * 4) {
      double d[4][4], acc[4][TS_NC];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int r = r4 + rr, k = lane + 32 * kk;
          d[rr][kk] = (r < jb && k <= r) ? D[(long long)r * NB + k] : 0.0;
        }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) acc[rr][c] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = lane + 32 * kk;
        if (k < jb) {
#pragma unroll
          for (int c = 0; c < TS_NC; ++c) {
            const double v = vec[off + k][c];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) acc[rr][c] = fma(d[rr][kk], v, acc[rr][c]);
          }
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) acc[rr][c] = warp_sum(acc[rr][c]);
      if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (r4 + rr < jb) {
#pragma unroll
            for (int c = 0; c < TS_NC; ++c) ybuf[r4 + rr][c] = acc[rr][c];
          }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < jb * TS_NC; idx += TS_THREADS) {
      const int r = idx / TS_NC, c = idx - r * TS_NC;
      if (c < ncol) W[(long long)(k0 + r) * ldw + col0 + c] = ybuf[r][c];
    }
    // rows below, inside the chunk: vec[i] -= L[i, k0 : k0 + jb] . y   (four rows per warp at a time)
    for (int i4 = off + jb + warp * 4; i4 < rows; i4 += (TS_THREADS / 32) * 4) {
      double l[4][4], acc[4][TS_NC];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int i = i4 + rr, k = lane + 32 * kk;
          l[rr][kk] = (i < rows && k < jb) ? L[(long long)(k_first + i) * lda + k0 + k] : 0.0;
        }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) acc[rr][c] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = lane + 32 * kk;
        if (k < jb) {
#pragma unroll
          for (int c = 0; c < TS_NC; ++c) {
            const double v = ybuf[k][c];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) acc[rr][c] = fma(l[rr][kk], v, acc[rr][c]);
          }
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) acc[rr][c] = warp_sum(acc[rr][c]);
      if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (i4 + rr < rows) {
#pragma unroll
            for (int c = 0; c < TS_NC; ++c) vec[i4 + rr][c] -= acc[rr][c];
          }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TS_THREADS) chunk_backward_kernel(
    const double* __restrict__ L, long long lda, long long stride_l, const double* __restrict__ dinv,
    long long dstride, const double* __restrict__ W, long long ldw, long long wstride,
    double* __restrict__ B, long long ldb, long long stride_b, int nrhs, int c0, int cend) {
  __shared__ double vec[TS_CHUNK][TS_NC];
  __shared__ double part[4][NB][TS_NC];
  __shared__ double xbuf[NB][TS_NC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = blockIdx.x * TS_NC;
  const int ncol = min(TS_NC, nrhs - col0);
  const long long b = blockIdx.y;
  L += b * stride_l; dinv += b * dstride; B += b * stride_b; W += b * wstride;
  const int rows = cend - c0;
  for (int idx = tid; idx < rows * TS_NC; idx += TS_THREADS) {
    const int r = idx / TS_NC, c = idx - r * TS_NC;
    vec[r][c] = c < ncol ? W[(long long)(c0 + r) * ldw + col0 + c] : 0.0;
  }
  __syncthreads();
  const int klast = ((cend - 1) / NB) * NB;
  for (int k0 = klast; k0 >= c0; k0 -= NB) {
    const int jb = min(NB, cend - k0);
    const int off = k0 - c0;
    const double* D = dinv + (long long)(k0 / NB) * NB * NB;
    // x = D^T * vec[off : off + jb]:  x[r] = sum_{k >= r} D[k][r] vec[off + k]; thread (q, r) takes
    // the rows k = q, q + 4, ... (coalesced along r), the four partial sums meet in shared memory
    {
      const int r = tid & (NB - 1), q = tid >> 7;
      double acc[TS_NC];
#pragma unroll
      for (int c = 0; c < TS_NC; ++c) acc[c] = 0.0;
      if (r < jb) {
        int k = r + ((q - r) & 3);                 // first k >= r with k % 4 == q
        for (; k < jb; k += 32) {                  // eight loads in flight per thread
          double d[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) d[u] = (k + 4 * u < jb) ? D[(long long)(k + 4 * u) * NB + r] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + 4 * u < jb) {
#pragma unroll
              for (int c = 0; c < TS_NC; ++c) acc[c] = fma(d[u], vec[off + k + 4 * u][c], acc[c]);
            }
        }
      }
#pragma unroll
      for (int c = 0; c < TS_NC; ++c) part[q][r][c] = acc[c];
    }
    __syncthreads();
    for (int idx = tid; idx < jb * TS_NC; idx += TS_THREADS) {
      const int r = idx / TS_NC, c = idx - r * TS_NC;
      const double x = (part[0][r][c] + part[1][r][c]) + (part[2][r][c] + part[3][r][c]);
      xbuf[r][c] = x;
      if (c < ncol) B[(long long)(k0 + r) * ldb + col0 + c] = x;
    }
    __syncthreads();
    // rows above, inside the chunk: vec[i] -= sum_k L[k0 + k, c0 + i] x[k]   (lanes along i)
    for (int t = warp; t * 32 < off; t += TS_THREADS / 32) {
      const int i = t * 32 + lane;
      if (i < off) {
        const double* Lcol = L + (long long)k0 * lda + c0 + i;
        double acc[TS_NC];
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) acc[c] = 0.0;
        for (int k = 0; k < jb; k += 8) {          // eight loads in flight per thread
          double l[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) l[u] = (k + u < jb) ? Lcol[(long long)(k + u) * lda] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k + u < jb) {
#pragma unroll
              for (int c = 0; c < TS_NC; ++c) acc[c] = fma(l[u], xbuf[k + u][c], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < TS_NC; ++c) vec[i][c] -= acc[c];
      }
    }
    __syncthreads();
  }
}

// phases: bit 0 = forward substitution for the 512-row chunks [chunk_lo, chunk_hi), bit 1 = backward
// substitution, bit 2 = clear the part of `work` the forward pass skips.
static int potrs_impl(const double* L, int N, long long lda, int batch, long long stride_l,
                      const double* dinv, double* B, int nrhs, long long ldb, long long stride_b,
                      int first_row, int lower_only, double* work, int phases, int chunk_lo,
                      int chunk_hi, cudaStream_t st) {
  PMR_REQUIRE(N >= 0 && nrhs >= 0 && batch >= 1 && batch <= 65535, "potrs: bad sizes");
  if (N == 0 || nrhs == 0) return 0;
  PMR_REQUIRE(L && dinv && B && work, "potrs: null pointer");
  PMR_REQUIRE(ldb >= nrhs, "potrs: ldb < nrhs");
  const long long dstride = pmr_potrf_dinv_doubles(N);
  const long long wstride = (long long)N * nrhs;
  const int nblk = (N + NB - 1) / NB;
  // forward: L Y = B, Y -> work.  Rows of B above first_row are zero on entry (columns of the
  // identity): their block rows are skipped, only the matching part of `work` is cleared.
  PMR_REQUIRE(first_row >= 0 && first_row <= N, "potrs: bad first_row");
  const int kb0 = first_row / NB;
  if (kb0 > 0 && (phases & 4))
    for (int b = 0; b < batch; ++b)
      PMR_CHECK_CUDA(cudaMemsetAsync(work + (long long)b * wstride, 0,
                                     sizeof(double) * (size_t)kb0 * NB * nrhs, st));
  // Both substitutions run in chunks of NBO rows: inside a chunk the 128-row blocks are solved one
  // after the other with small in-chunk updates, then ONE rank-NBO update carries the chunk into
  // the rest of the right-hand side (k = 512 keeps the DMMA pipeline full; k = 128 did not).
  constexpr int NBO = 512;
  const int cb0 = std::max((kb0 * NB) / NBO, chunk_lo);
  const int cb1 = std::min((N + NBO - 1) / NBO, chunk_hi);
  for (int cb = cb0; cb < cb1 && (phases & 1); ++cb) {
    const int c0 = cb * NBO;
    const int cend = std::min(N, c0 + NBO);
    if (g_fused_chunks) {
      dim3 grid((unsigned)((nrhs + TS_NC - 1) / TS_NC), (unsigned)batch);
      chunk_forward_kernel<<<grid, TS_THREADS, 0, st>>>(L, lda, stride_l, dinv, dstride, B, ldb,
                                                        stride_b, work, nrhs, wstride, nrhs,
                                                        std::max(c0, kb0 * NB), cend);
      PMR_CHECK_LAUNCH("chunk_forward_kernel");
    }
    for (int k0 = std::max(c0, kb0 * NB); k0 < cend && !g_fused_chunks; k0 += NB) {
      const int jb = std::min(NB, N - k0), inrest = cend - k0 - jb;
      const double* dk = dinv + (long long)(k0 / NB) * NB * NB;
      pmr_gemm_t g = make_gemm(jb, nrhs, jb, 1.0, dk, NB, 1, B + (long long)k0 * ldb, ldb, 1, 0.0,
                               work + (long long)k0 * nrhs, nrhs, 1);
      g.batch1 = batch; g.a_b1 = dstride; g.b_b1 = stride_b; g.c_b1 = wstride;
      PMR_TRY(gemm(g, st));
      if (inrest > 0) {
        pmr_gemm_t u = make_gemm(inrest, nrhs, jb, -1.0, L + (long long)(k0 + jb) * lda + k0, lda, 1,
                                 work + (long long)k0 * nrhs, nrhs, 1, 1.0,
                                 B + (long long)(k0 + jb) * ldb, ldb, 1);
        u.batch1 = batch; u.a_b1 = stride_l; u.b_b1 = wstride; u.c_b1 = stride_b;
        PMR_TRY(gemm(u, st));
      }
    }
    if (cend < N) {
      const int kfirst = std::max(c0, kb0 * NB);   // rows of Y above it are zero
      pmr_gemm_t u = make_gemm(N - cend, nrhs, cend - kfirst, -1.0,
                               L + (long long)cend * lda + kfirst, lda, 1,
                               work + (long long)kfirst * nrhs, nrhs, 1, 1.0,
                               B + (long long)cend * ldb, ldb, 1);
      u.batch1 = batch; u.a_b1 = stride_l; u.b_b1 = wstride; u.c_b1 = stride_b;
      PMR_TRY(gemm(u, st));
    }
  }
  // backward: L^T X = Y, X -> B.  With lower_only the rows above first_row (rounded down to a
  // block) are not wanted (a symmetric result whose upper part nobody reads) and stay zero.
  const int r0 = lower_only ? kb0 * NB : 0;
  const int nchunk = (N + NBO - 1) / NBO;
  for (int cb = nchunk - 1; cb >= r0 / NBO && (phases & 2); --cb) {
    const int c0 = std::max(cb * NBO, r0), cend = std::min(N, cb * NBO + NBO);
    const int klast = ((cend - 1) / NB) * NB;
    if (g_fused_chunks) {
      dim3 grid((unsigned)((nrhs + TS_NC - 1) / TS_NC), (unsigned)batch);
      chunk_backward_kernel<<<grid, TS_THREADS, 0, st>>>(L, lda, stride_l, dinv, dstride, work, nrhs,
                                                         wstride, B, ldb, stride_b, nrhs, c0, cend);
      PMR_CHECK_LAUNCH("chunk_backward_kernel");
    }
    for (int k0 = klast; k0 >= c0 && !g_fused_chunks; k0 -= NB) {
      const int jb = std::min(NB, N - k0);
      const double* dk = dinv + (long long)(k0 / NB) * NB * NB;
      pmr_gemm_t g = make_gemm(jb, nrhs, jb, 1.0, dk, 1, NB, work + (long long)k0 * nrhs, nrhs, 1,
                               0.0, B + (long long)k0 * ldb, ldb, 1);
      g.batch1 = batch; g.a_b1 = dstride; g.b_b1 = wstride; g.c_b1 = stride_b;
      PMR_TRY(gemm(g, st));
      if (k0 > c0) {
        pmr_gemm_t u = make_gemm(k0 - c0, nrhs, jb, -1.0, L + (long long)k0 * lda + c0, 1, lda,
                                 B + (long long)k0 * ldb, ldb, 1, 1.0,
                                 work + (long long)c0 * nrhs, nrhs, 1);
        u.batch1 = batch; u.a_b1 = stride_l; u.b_b1 = stride_b; u.c_b1 = wstride;
        PMR_TRY(gemm(u, st));
      }
    }
    if (c0 > r0) {
      pmr_gemm_t u = make_gemm(c0 - r0, nrhs, cend - c0, -1.0, L + (long long)c0 * lda + r0, 1, lda,
                               B + (long long)c0 * ldb, ldb, 1, 1.0,
                               work + (long long)r0 * nrhs, nrhs, 1);
      u.batch1 = batch; u.a_b1 = stride_l; u.b_b1 = stride_b; u.c_b1 = wstride;
      PMR_TRY(gemm(u, st));
    }
  }
  return 0;
}

extern "C" void pmr_set_fused_substitution(int on) { g_fused_chunks = on != 0; }

extern "C" int pmr_potrs(const double* L, int N, long long lda, int batch, long long stride_l,
                         const double* dinv, double* B, int nrhs, long long ldb,
                         long long stride_b, int first_row, int lower_only, double* work,
                         void* stream) {
  return potrs_impl(L, N, lda, batch, stride_l, dinv, B, nrhs, ldb, stride_b, first_row, lower_only,
                    work, 1 | 2 | 4, 0, 1 << 30, as_stream(stream));
}

// phases: 1 = forward only (B -> work), 2 = backward only (work -> B), 3 = both; batched like pmr_potrs
extern "C" int pmr_potrs_ex(const double* L, int N, long long lda, int batch, long long stride_l,
                            const double* dinv, double* B, int nrhs, long long ldb,
                            long long stride_b, double* work, int phases, void* stream) {
  PMR_REQUIRE(phases >= 1 && phases <= 3, "potrs_ex: phases must be 1, 2 or 3");
  return potrs_impl(L, N, lda, batch, stride_l, dinv, B, nrhs, ldb, stride_b, 0, 0, work,
                    phases | 4, 0, 1 << 30, as_stream(stream));
}

extern "C" int pmr_potrs_phase(const double* L, int N, long long lda, const double* dinv, double* B,
                               int nrhs, long long ldb, int first_row, int lower_only, double* work,
                               int backward, int chunk_lo, int chunk_hi, void* stream) {
  return potrs_impl(L, N, lda, 1, 0, dinv, B, nrhs, ldb, 0, first_row, lower_only, work,
                    backward ? 2 : 1, chunk_lo, chunk_hi, as_stream(stream));
}

// Z = inv(L), L lower triangular with the inverses of its 128 x 128 diagonal blocks in `dinv`:
// recursive doubling.  Level 0 places the diagonal-block inverses; at block size b = 128 * 2^l every
// pair of neighbouring blocks [[L11, 0], [L21, L22]] gets  Z21 = -Z22 L21 Z11  from two GEMMs that are
// BATCHED over all pairs of the level (log2(N/128) levels, about 2 launches each, against one
// latency-bound substitution step per 128 columns: 56 steps at N = 7168).  Triangular operands skip
// their zero halves through the k-range flags (T^T = Z11^T L21^T with k >= row tile; Z21 = -Z22 T with
// k <= row tile), so the work is the N^3/3 of a triangular inversion.  T^T lives in the (dead) Z12
// block; entries of Z above its 128-wide block diagonal are scratch on exit (the Gram product below
// starts its k loop at the row tile and never reads them), the block diagonal itself is exact.
extern "C" int pmr_trtri_tree(const double* L, int N, long long lda, const double* dinv, double* Z,
                              long long ldz, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0, "trtri_tree: bad N");
  if (N == 0) return 0;
  PMR_REQUIRE(L && dinv && Z, "trtri_tree: null pointer");
  PMR_REQUIRE(lda >= N && ldz >= N, "trtri_tree: leading dimension too small");
  const int nblk = (N + NB - 1) / NB;
  {
    dim3 grid(nblk, NB);
    place_diag_blocks_kernel<<<grid, 128, 0, st>>>(dinv, Z, ldz, N);
    PMR_CHECK_LAUNCH("place_diag_blocks_kernel");
  }
  for (long long b = NB; b < N; b *= 2) {
    const long long span = 2 * b;                  // rows covered by one pair
    const int npairs = (int)((N + span - 1) / span);
    // pairs [0, nfull) have a complete second block; at most one ragged pair follows
    int nfull = 0;
    while (nfull < npairs && (long long)nfull * span + span <= N) ++nfull;
    for (int part = 0; part < 2; ++part) {
      const int first = part == 0 ? 0 : nfull;
      const int count = part == 0 ? nfull : npairs - nfull;
      if (count <= 0) continue;
      const long long r0 = (long long)first * span;
      const long long n1 = b;
      const long long n2 = part == 0 ? b : (long long)N - (r0 + b);
      if (n2 <= 0) continue;                       // a lone block at the end: nothing to couple
      const double* L21 = L + (r0 + b) * lda + r0;
      double* Z11 = Z + r0 * ldz + r0;
      double* Z12 = Z + r0 * ldz + (r0 + b);       // scratch: T^T (n1 x n2)
      double* Z21 = Z + (r0 + b) * ldz + r0;
      double* Z22 = Z + (r0 + b) * ldz + (r0 + b);
      const long long ps_l = span * (lda + 1), ps_z = span * (ldz + 1);
      // T^T(c, r) = sum_{k >= c} Z11(k, c) L21(r, k)
      pmr_gemm_t g1 = make_gemm((int)n1, (int)n2, (int)n1, 1.0, Z11, 1, ldz, L21, 1, lda, 0.0, Z12, ldz, 1);
      g1.flags = PMR_GEMM_KSTART_M;
      g1.batch1 = count; g1.a_b1 = ps_z; g1.b_b1 = ps_l; g1.c_b1 = ps_z;
      PMR_TRY(gemm(g1, st));
      // Z21(r, c) = -sum_{k <= r} Z22(r, k) T(k, c),  T(k, c) = T^T(c, k)
      pmr_gemm_t g2 = make_gemm((int)n2, (int)n1, (int)n2, -1.0, Z22, ldz, 1, Z12, 1, ldz, 0.0, Z21, ldz, 1);
      g2.flags = PMR_GEMM_KEND_M;
      g2.batch1 = count; g2.a_b1 = ps_z; g2.b_b1 = ps_z; g2.c_b1 = ps_z;
      PMR_TRY(gemm(g2, st));
    }
  }
  return 0;
}

// out[c0 + r', c0 + c'] (r' >= c', c' < ncols) = sum_k Z(k, c0 + r') Z(k, c0 + c'): the columns
// [c0, c0 + ncols) of the lower triangle of Z^T Z for lower-triangular Z (rows >= c0 only: the
// sub-block Z[c0:, c0:] is itself lower triangular, so nothing above it contributes).
extern "C" int pmr_gram_lower_cols(const double* Z, int N, long long ldz, int c0, int ncols,
                                   double* out, long long ldo, void* stream) {
  PMR_REQUIRE(N >= 0 && c0 >= 0 && ncols >= 0 && c0 + ncols <= N, "gram_lower_cols: bad range");
  if (N == 0 || ncols == 0) return 0;
  PMR_REQUIRE(Z && out, "gram_lower_cols: null pointer");
  PMR_REQUIRE(c0 % NB == 0, "gram_lower_cols: c0 must be a multiple of 128 (the k loop starts at the row tile)");
  const int m = N - c0;
  const double* Zs = Z + (long long)c0 * ldz + c0;
  pmr_gemm_t s = make_gemm(m, ncols, m, 1.0, Zs, 1, ldz, Zs, ldz, 1, 0.0,
                           out + (long long)c0 * ldo + c0, ldo, 1);
  s.flags = PMR_GEMM_LOWER | PMR_GEMM_KSTART_M;
  return gemm(s, as_stream(stream));
}

extern "C" int pmr_potri_lower(const double* L, int N, long long lda, const double* dinv,
                               double* out, long long ldo, double* work, void* stream) {
  PMR_REQUIRE(N >= 0, "potri: bad N");
  if (N == 0) return 0;
  PMR_REQUIRE(L && dinv && out && work, "potri: null pointer");
  PMR_REQUIRE(ldo >= N, "potri: ldo < N");
  // Z = inv(L) -> work (recursive doubling), then out(lower) = Z^T Z
  PMR_TRY(pmr_trtri_tree(L, N, lda, dinv, work, N, stream));
  return pmr_gram_lower_cols(work, N, N, 0, N, out, ldo, stream);
}

extern "C" int pmr_symmetrize_lower(double* A, int N, long long lda, void* stream) {
  PMR_REQUIRE(N >= 0 && lda >= N, "symmetrize_lower: bad N/lda");
  if (N == 0) return 0;
  PMR_REQUIRE(A != nullptr, "symmetrize_lower: null pointer");
  const long long nt = (N + 31) / 32;
  const long long blocks = nt * (nt + 1) / 2;
  PMR_REQUIRE(blocks < (1LL << 31), "symmetrize_lower: grid too large");
  symmetrize_lower_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(A, N, lda);
  PMR_CHECK_LAUNCH("symmetrize_lower_kernel");
  return 0;
}

extern "C" int pmr_shifted_combine(const double* M, const double* W, int N, long long ldm,
                                   long long ldw, const double* shift_a_host,
                                   const double* shift_b_host, int nf, double* S, long long lds,
                                   long long stride_s, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0 && nf >= 0, "shifted_combine: bad sizes");
  if (N == 0 || nf == 0) return 0;
  PMR_REQUIRE(M && S && shift_a_host && shift_b_host, "shifted_combine: null pointer");
  dim3 grid((N + 255) / 256, N);
  for (int f0 = 0; f0 < nf; f0 += SHIFT_MAX) {
    const int cnt = std::min(SHIFT_MAX, nf - f0);
    ShiftPack sh{};
    for (int f = 0; f < cnt; ++f) { sh.a[f] = shift_a_host[f0 + f]; sh.b[f] = shift_b_host[f0 + f]; }
    shifted_combine_kernel<<<grid, 256, 0, st>>>(M, W, N, ldm, ldw, cnt, S + (long long)f0 * stride_s,
                                                 lds, stride_s, sh);
    PMR_CHECK_LAUNCH("shifted_combine_kernel");
  }
  return 0;
}

// --------------------------------------------------------------------------------------- LU
extern "C" long long pmr_getrf_work_doubles(int N) {
  (void)N;
  return 2 * 1024;  // pivot-search partials (values + indices packed as doubles)
}

extern "C" int pmr_getrf(double* A, int N, long long lda, int* ipiv, int* info, double* work,
                         void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0, "getrf: bad N");
  PMR_REQUIRE(A && ipiv && info && work, "getrf: null pointer");
  PMR_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (N == 0) return 0;
  double* pval = work;
  int* pidx = reinterpret_cast<int*>(work + 1024);
  for (int k0 = 0; k0 < N; k0 += PB) {
    const int pb = std::min(PB, N - k0);
    const int pend = k0 + pb;
    for (int j = k0; j < pend; ++j) {
      const int rows = N - j;
      const int nparts = std::max(1, std::min(512, (rows + 1023) / 1024));
      lu_pivot_partial_kernel<<<nparts, 256, 0, st>>>(A, lda, N, j, pval, pidx);
      PMR_CHECK_LAUNCH("lu_pivot_partial_kernel");
      lu_pivot_swap_kernel<<<1, 256, 0, st>>>(A, lda, N, j, pval, pidx, nparts, ipiv, info);
      PMR_CHECK_LAUNCH("lu_pivot_swap_kernel");
      if (rows > 1) {
        const int wpb = 8;
        lu_panel_update_kernel<<<(rows - 1 + wpb - 1) / wpb, wpb * 32, 0, st>>>(A, lda, N, j, pend);
        PMR_CHECK_LAUNCH("lu_panel_update_kernel");
      }
    }
    const int rest = N - pend;
    if (rest > 0) {
      lu_trsm_u12_kernel<<<(rest + 127) / 128, 128, 0, st>>>(A, lda, k0, pb, N);
      PMR_CHECK_LAUNCH("lu_trsm_u12_kernel");
      pmr_gemm_t g = make_gemm(rest, rest, pb, -1.0, A + (long long)pend * lda + k0, lda, 1,
                               A + (long long)k0 * lda + pend, lda, 1, 1.0,
                               A + (long long)pend * lda + pend, lda, 1);
      PMR_TRY(gemm(g, st));
    }
  }
  return 0;
}

extern "C" int pmr_getrs(const double* LU, int N, long long lda, const int* ipiv, double* B,
                         int nrhs, long long ldb, double* work, void* stream) {
  cudaStream_t st = as_stream(stream);
  (void)work;
  PMR_REQUIRE(N >= 0 && nrhs >= 0, "getrs: bad sizes");
  if (N == 0 || nrhs == 0) return 0;
  PMR_REQUIRE(LU && ipiv && B, "getrs: null pointer");
  lu_apply_pivots_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(B, ldb, nrhs, N, ipiv);
  PMR_CHECK_LAUNCH("lu_apply_pivots_kernel");
  // forward (unit lower), right-looking
  for (int k0 = 0; k0 < N; k0 += PB) {
    const int pb = std::min(PB, N - k0), rest = N - k0 - pb;
    lu_diag_solve_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(LU, lda, k0, pb, B, ldb, nrhs, 0);
    PMR_CHECK_LAUNCH("lu_diag_solve_kernel");
    if (rest > 0) {
      pmr_gemm_t g = make_gemm(rest, nrhs, pb, -1.0, LU + (long long)(k0 + pb) * lda + k0, lda, 1,
                               B + (long long)k0 * ldb, ldb, 1, 1.0,
                               B + (long long)(k0 + pb) * ldb, ldb, 1);
      PMR_TRY(gemm(g, st));
    }
  }
  // backward (upper)
  const int nblk = (N + PB - 1) / PB;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int k0 = kb * PB, pb = std::min(PB, N - k0);
    lu_diag_solve_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(LU, lda, k0, pb, B, ldb, nrhs, 1);
    PMR_CHECK_LAUNCH("lu_diag_solve_kernel");
    if (k0 > 0) {
      pmr_gemm_t g = make_gemm(k0, nrhs, pb, -1.0, LU + k0, lda, 1, B + (long long)k0 * ldb, ldb,
                               1, 1.0, B, ldb, 1);
      PMR_TRY(gemm(g, st));
    }
  }
  return 0;
}
