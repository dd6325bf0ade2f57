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
#include <cooperative_groups.h>

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

// One launch per panel column j (rows [j, N), a warp per row):
//   1. unless j is the panel's first column, finish column j-1: l = A[i][j-1] / pivot, stored in place, and
//      A[i][j..pend) -= l * A[j-1][j..pend)  -- the rank-1 step restricted to the panel;
//   2. unless j == pend (the call that only finishes the panel's last column), search column j for its pivot:
//      per-block candidates, then the last block to arrive (ticket counter) picks the pivot, records it
//      and swaps rows j <-> p inside the panel columns [k0, pend).  The interchange of the columns outside
//      the panel is applied once per panel by lu_laswp_kernel (LAPACK's order of operations).
// Ties go to the smallest row index, as in LAPACK's idamax.
constexpr int LU_WARPS = 8;
constexpr int LU_MAX_PARTS = 512;

__global__ void __launch_bounds__(LU_WARPS * 32)
lu_column_kernel(double* A, long long lda, int N, int k0, int pend, int j, double* pval, int* pidx,
                 int* counter, int* ipiv, int* info) {
  __shared__ double prow[PB];
  __shared__ double sv[LU_WARPS * 32];
  __shared__ int si[LU_WARPS * 32];
  __shared__ int s_last, s_piv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool finish_prev = j > k0;
  const bool search = j < pend;
  double piv = 1.0;
  if (finish_prev) {
    for (int c = threadIdx.x; c < pend - j; c += blockDim.x) prow[c] = A[(long long)(j - 1) * lda + j + c];
    piv = A[(long long)(j - 1) * lda + (j - 1)];
  }
  __syncthreads();
  double best = -1.0;
  int bi = j;
  for (int i = j + blockIdx.x * LU_WARPS + warp; i < N; i += gridDim.x * LU_WARPS) {
    double* row = A + (long long)i * lda;
    double cand = 0.0;
    if (finish_prev) {
      const double l = row[j - 1] / piv;
      __syncwarp();
      for (int c = j + lane; c < pend; c += 32) {
        const double v = row[c] - l * prow[c - j];
        row[c] = v;
        if (c == j) cand = v;
      }
      if (lane == 0) row[j - 1] = l;
    } else if (lane == 0) {
      cand = row[j];
    }
    if (search && lane == 0) {
      const double v = fabs(cand);
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
  }
  if (!search) return;
  // block candidate (lane 0 of every warp holds one)
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = sv[0];
    int ix = si[0];
    for (int w = 1; w < LU_WARPS; ++w)
      if (sv[w] > b || (sv[w] == b && si[w] < ix)) { b = sv[w]; ix = si[w]; }
    pval[blockIdx.x] = b;
    pidx[blockIdx.x] = ix;
    __threadfence();
    const int ticket = atomicAdd(counter, 1);
    s_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last block: reduce the per-block candidates
  {
    double b = -1.0;
    int ix = j;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
      const double v = __ldcg(pval + q);
      const int i2 = __ldcg(pidx + q);
      if (v > b || (v == b && i2 < ix)) { b = v; ix = i2; }
    }
    sv[threadIdx.x] = b;
    si[threadIdx.x] = ix;
  }
  __syncthreads();
  for (int s = (LU_WARPS * 32) >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const double v = sv[threadIdx.x + s];
      const int i2 = si[threadIdx.x + s];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ipiv[j] = si[0];
    if (!(sv[0] > 0.0) && *info == 0) *info = j + 1;
    s_piv = si[0];
    *counter = 0;
  }
  __syncthreads();
  const int p = s_piv;
  if (p == j) return;
  for (int c = k0 + threadIdx.x; c < pend; c += blockDim.x) {
    const double a = __ldcg(A + (long long)j * lda + c);
    const double b = __ldcg(A + (long long)p * lda + c);
    A[(long long)j * lda + c] = b;
    A[(long long)p * lda + c] = a;
  }
}

// Row interchanges ipiv[p_lo..p_hi) applied to the columns [a_lo, a_hi) and [b_lo, b_hi), one after the
// other by a thread per column (the fallback when the row range does not fit lu_perm_build_kernel).
__global__ void lu_laswp_kernel(double* __restrict__ A, long long lda, int p_lo, int p_hi,
                                const int* __restrict__ ipiv, int a_lo, int a_hi, int b_lo, int b_hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int na = a_hi - a_lo;
  if (t >= na + (b_hi - b_lo)) return;
  const int c = t < na ? a_lo + t : b_lo + (t - na);
  for (int j = p_lo; j < p_hi; ++j) {
    const int p = ipiv[j];
    if (p != j) {
      const double a = A[(long long)j * lda + c];
      A[(long long)j * lda + c] = A[(long long)p * lda + c];
      A[(long long)p * lda + c] = a;
    }
  }
}

// The same interchanges as ONE permutation: rows [p_lo, N) tracked in shared memory (perm[t] = the row
// that ends at position t), the swaps replayed by one thread, then the list of positions that move
// (at most 2 (p_hi - p_lo)) written out as (target, source) pairs for the gather / scatter pair below.
constexpr int LU_OB = 256;  // outer block of the factorisation = the most interchanges replayed at once
constexpr int LU_PERM_MAX_ROWS = 51200;

__global__ void __launch_bounds__(256)
lu_perm_build_kernel(const int* __restrict__ ipiv, int p_lo, int p_hi, int N, int* __restrict__ list_t,
                     int* __restrict__ list_src, int* __restrict__ count) {
  extern __shared__ int lu_perm[];
  __shared__ int piv[LU_OB];
  __shared__ int cnt;
  const int n = N - p_lo, ns = p_hi - p_lo;
  for (int i = threadIdx.x; i < n; i += blockDim.x) lu_perm[i] = i;
  for (int q = threadIdx.x; q < ns; q += blockDim.x) piv[q] = ipiv[p_lo + q] - p_lo;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < ns; ++q) {
      const int p = piv[q];
      if (p != q) {
        const int t = lu_perm[q];
        lu_perm[q] = lu_perm[p];
        lu_perm[p] = t;
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < ns; q += blockDim.x) {
    if (lu_perm[q] != q) {
      const int m = atomicAdd(&cnt, 1);
      list_t[m] = p_lo + q;
      list_src[m] = p_lo + lu_perm[q];
    }
    const int p = piv[q];
    if (p >= ns && lu_perm[p] != p) {
      bool first = true;
      for (int q2 = 0; q2 < q; ++q2) first = first && piv[q2] != p;
      if (first) {
        const int m = atomicAdd(&cnt, 1);
        list_t[m] = p_lo + p;
        list_src[m] = p_lo + lu_perm[p];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *count = cnt;
}

// tmp[m][.] <- A[src_m][cols]   (SCATTER = false)      A[t_m][cols] <- tmp[m][.]   (SCATTER = true)
template <bool SCATTER>
__global__ void lu_perm_move_kernel(double* __restrict__ A, long long lda, const int* __restrict__ list_t,
                                    const int* __restrict__ list_src, const int* __restrict__ count,
                                    double* __restrict__ tmp, int a_lo, int a_hi, int b_lo, int b_hi) {
  const int m = blockIdx.y;
  if (m >= *count) return;
  const int na = a_hi - a_lo, ncols = na + (b_hi - b_lo);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncols) return;
  const int c = t < na ? a_lo + t : b_lo + (t - na);
  if (SCATTER) A[(long long)list_t[m] * lda + c] = tmp[(long long)m * ncols + t];
  else tmp[(long long)m * ncols + t] = A[(long long)list_src[m] * lda + c];
}

// Whole pivot sequence of a factorisation as one row permutation (for the right-hand sides of getrs).
__global__ void __launch_bounds__(1024)
lu_perm_full_kernel(const int* __restrict__ ipiv, int N, int* __restrict__ perm_out) {
  extern __shared__ int lu_perm[];
  for (int i = threadIdx.x; i < N; i += blockDim.x) lu_perm[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < N; ++q) {
      const int p = ipiv[q];
      if (p != q) {
        const int t = lu_perm[q];
        lu_perm[q] = lu_perm[p];
        lu_perm[p] = t;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) perm_out[i] = lu_perm[i];
}

// out[i][c] = B[perm[i]][c]  (perm == nullptr: plain copy back)
__global__ void lu_permute_rows_kernel(const double* __restrict__ src, long long lds,
                                       double* __restrict__ dst, long long ldd, int N, int nrhs,
                                       const int* __restrict__ perm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (c >= nrhs || i >= N) return;
  const int r = perm ? perm[i] : i;
  dst[(long long)i * ldd + c] = src[(long long)r * lds + c];
}

// ---- inner panel on one thread-block cluster ---------------------------------------------------
// LUC_W columns of the outer panel, rows [j0, N), factorised by ONE launch: the LUC_CL CTAs of a cluster
// hold the rows in registers (rpc rows per CTA, staged through shared memory) and run the LUC_W column steps with a
// single cluster barrier per column.  Before the barrier every CTA writes its pivot candidate (value,
// row index and the candidate row's LUC_W entries) into a slot of every CTA of the cluster through
// distributed shared memory, and CTA 0 writes row j the same way; after it each CTA knows the pivot, has
// the pivot row for the rank-1 step, and the owners of rows j and p put the other row's entries in place
// -- no second exchange.  Slots are double-buffered by column parity (a CTA can be at most one barrier
// ahead).  The interchange of the outer panel's other columns is done in global memory by fixed threads
// of CTA 0 (one per column, so a column's swaps stay in program order), the stores of column j delayed to
// column j+1 so the load latency is off the critical path.  The same threads finish with
// U12 <- inv(L11) A12 for the outer panel's columns right of the inner panel.
namespace cg = cooperative_groups;
constexpr int LUC_W = 8;
constexpr int LUC_CL = 8;
constexpr int LUC_THREADS = 512;
constexpr int LUC_RPC_MAX = 2048;  // rows per CTA: four register-resident rows per thread
constexpr int LUC_PAD = 4;         // column stride rpc + 4: the row-wise load / store spreads over the banks
static_assert(LUC_W == 8 && LUC_CL == 8, "row-wise staging and the slot writes assume 8 columns, 8 CTAs");

struct LucSlot {
  unsigned long long key;  // luc_key of the candidate's magnitude
  int idx;
  int pad;
  double row[LUC_W];
};

__device__ __forceinline__ bool luc_better(double v, int i, double bv, int bi) {
  return v > bv || (v == bv && i < bi);
}

// Pivot candidates as ordered integers: a non-negative double compares like its bit pattern, +1 keeps
// zero apart from "no candidate" (key 0).  Warp-wide arg-max with the smallest index on ties in three
// redux operations (high word, low word, index); every lane gets the result.
__device__ __forceinline__ unsigned long long luc_key(double v) {
  return v >= 0.0 ? (unsigned long long)__double_as_longlong(v) + 1ull : 0ull;
}
__device__ __forceinline__ double luc_value(unsigned long long key) {
  return key ? __longlong_as_double((long long)(key - 1ull)) : -1.0;
}
__device__ __forceinline__ void luc_warp_argmax(unsigned long long& key, int& idx) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  bool cand = hi == mh;
  const unsigned ml = __reduce_max_sync(0xffffffffu, cand ? lo : 0u);
  cand = cand && lo == ml;
  idx = __reduce_min_sync(0xffffffffu, cand ? idx : 0x7fffffff);
  key = ((unsigned long long)mh << 32) | ml;
}

// Rows live in registers (RPT rows of LUC_W entries per thread: thread tid owns the local rows tid + k *
// LUC_THREADS), so the rank-1 steps and the candidate scans run at register speed; shared memory only
// stages the panel in and out (eight lanes per 64-byte row) and carries the pivot exchange.
template <int RPT>
__global__ void __cluster_dims__(LUC_CL, 1, 1) __launch_bounds__(LUC_THREADS)
lu_inner_panel_kernel(double* A, long long lda, int N, int k0, int pend, int j0, int w, int rpc,
                      int* ipiv, int* info) {
  extern __shared__ double luc_a[];  // [LUC_W][rpc + LUC_PAD], staging only
  __shared__ LucSlot slots[2][LUC_CL];
  __shared__ double rowj[2][LUC_W];
  __shared__ unsigned long long wk[32];
  __shared__ int wi[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nrows = N - j0;
  const int base = rank * rpc;
  const int myrows = max(0, min(rpc, nrows - base));
  const int ld = rpc + LUC_PAD;
  double* a = luc_a;
  // development aid (pmr_debug_potf2_timing): phase stamps of the first inner panel of a factorisation
#define LUC_TICK(slot)                                                                   \
  do {                                                                                   \
    if (tid == 0 && rank == 0 && j0 == 0 && g_potf2_clock) g_potf2_clock[slot] = clock64(); \
  } while (0)
  LUC_TICK(0);

  for (int e = tid; e < myrows * LUC_W; e += LUC_THREADS) {
    const int li = e >> 3, c = e & 7;
    if (c < w) a[c * ld + li] = A[(long long)(j0 + base + li) * lda + j0 + c];
  }
  __syncthreads();
  double x[RPT][LUC_W];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int li = tid + k * LUC_THREADS;
#pragma unroll
    for (int c = 0; c < LUC_W; ++c) x[k][c] = (li < myrows && c < w) ? a[c * ld + li] : 0.0;
  }
  LUC_TICK(1);

  // the thread of CTA 0 that owns outer-panel column sc (outside the inner panel)
  const int sc = k0 + tid;
  const bool swapper = rank == 0 && sc < pend && (sc < j0 || sc >= j0 + w);
  bool pending = false;
  double va = 0.0, vb = 0.0;
  long long pa = 0, pb2 = 0;

#pragma unroll
  for (int jj = 0; jj < LUC_W; ++jj) {
    if (jj < w) {  // uniform over the cluster
      const int par = jj & 1;
      // pivot candidate of this thread, warp, CTA
      double best = -1.0;
      int bi = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int li = tid + k * LUC_THREADS;
        if (li < myrows && base + li >= jj) {
          const double v = fabs(x[k][jj]);
          if (luc_better(v, base + li, best, bi)) { best = v; bi = base + li; }
        }
      }
      unsigned long long key = luc_key(best);
      luc_warp_argmax(key, bi);
      if (lane == 0) { wk[warp] = key; wi[warp] = bi; }
      __syncthreads();
      // every warp finishes the CTA-level reduction for itself (no second block barrier)
      key = lane < LUC_THREADS / 32 ? wk[lane] : 0ull;
      bi = lane < LUC_THREADS / 32 ? wi[lane] : 0x7fffffff;
      luc_warp_argmax(key, bi);
      LUC_TICK(2 + 5 * jj);
      // the warp that owns the CTA's candidate row writes it into this CTA's slot of every CTA: the
      // owner lane hands entry c to lane c, then lane 8 dd + c stores entry c at destination 4 r + dd
      const int lo = bi - base;
      const int owner_tid = lo & (LUC_THREADS - 1);
      if (key != 0ull && (owner_tid >> 5) == warp) {  // uniform over the warp
        const int kk = lo / LUC_THREADS;
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < LUC_W; ++c) {
          double v = x[0][c];
#pragma unroll
          for (int k = 1; k < RPT; ++k)
            if (kk == k) v = x[k][c];
          v = __shfl_sync(0xffffffffu, v, owner_tid & 31);
          if (lane == c) mine = v;
        }
        const int c = lane & 7;
        const double val = __shfl_sync(0xffffffffu, mine, c);
#pragma unroll
        for (int r = 0; r < LUC_CL / 4; ++r) {
          LucSlot* rs = cluster.map_shared_rank(&slots[par][rank], r * 4 + (lane >> 3));
          rs->row[c] = val;
          if (c == 0) { rs->key = key; rs->idx = bi; }
        }
      } else if (key == 0ull && tid < LUC_CL) {
        LucSlot* rs = cluster.map_shared_rank(&slots[par][rank], tid);
        rs->key = 0ull;
        rs->idx = 0x7fffffff;
      }
      if (rank == 0 && warp == 0) {  // row j is local row jj of CTA 0: lane jj of warp 0, k = 0
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < LUC_W; ++c) {
          const double v = __shfl_sync(0xffffffffu, x[0][c], jj);
          if (lane == c) mine = v;
        }
        const int c = lane & 7;
        const double val = __shfl_sync(0xffffffffu, mine, c);
#pragma unroll
        for (int r = 0; r < LUC_CL / 4; ++r) {
          double* rr = cluster.map_shared_rank(&rowj[par][0], r * 4 + (lane >> 3));
          rr[c] = val;
        }
      }
      LUC_TICK(3 + 5 * jj);
      cluster.sync();
      LUC_TICK(4 + 5 * jj);

      // the pivot: largest candidate, smallest row index on ties; none valid -> keep row j
      unsigned long long pk = 0ull;
      int pidx = 0x7fffffff, wr = 0;
#pragma unroll
      for (int d = 0; d < LUC_CL; ++d) {
        const unsigned long long kd = slots[par][d].key;
        const int ix = slots[par][d].idx;
        if (kd > pk || (kd == pk && ix < pidx)) { pk = kd; pidx = ix; wr = d; }
      }
      if (pk == 0ull) pidx = jj;
      const double* prow_s = (pidx == jj) ? rowj[par] : slots[par][wr].row;
      double prow[LUC_W];
#pragma unroll
      for (int c = 0; c < LUC_W; ++c) prow[c] = prow_s[c];
      const double piv = prow[jj];
      // multipliers by the reciprocal pivot, as LAPACK's dgetf2 / dgetrf2 scale the column (guarded
      // against a reciprocal that would overflow)
      const bool use_rpiv = fabs(piv) >= 2.2250738585072014e-308;
      const double rpiv = 1.0 / piv;
      if (rank == 0 && tid == 0) {
        ipiv[j0 + jj] = j0 + pidx;
        if (pk <= 1ull && *info == 0) *info = j0 + jj + 1;  // key 1 is the value 0.0
      }
      if (pidx != jj) {
        // row j's entries go to the pivot row's position, the pivot row to position j
        const int lo = pidx - base;
        if (lo >= 0 && lo < myrows && (lo & (LUC_THREADS - 1)) == tid) {
          const int kk = lo / LUC_THREADS;
#pragma unroll
          for (int c = 0; c < LUC_W; ++c) {
            const double v = rowj[par][c];
#pragma unroll
            for (int k = 0; k < RPT; ++k)
              if (kk == k) x[k][c] = v;
          }
        }
        if (rank == 0 && tid == jj) {
#pragma unroll
          for (int c = 0; c < LUC_W; ++c) x[0][c] = prow[c];
        }
      }
      if (swapper) {
        if (pending) { A[pa] = vb; A[pb2] = va; pending = false; }
        if (pidx != jj) {
          pa = (long long)(j0 + jj) * lda + sc;
          pb2 = (long long)(j0 + pidx) * lda + sc;
          va = A[pa];
          vb = A[pb2];
          pending = true;
        }
      }
      LUC_TICK(5 + 5 * jj);
      // rank-1 step on the rows below j
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int li = tid + k * LUC_THREADS;
        if (li < myrows && base + li > jj) {
          const double l = use_rpiv ? x[k][jj] * rpiv : x[k][jj] / piv;
          x[k][jj] = l;
#pragma unroll
          for (int c = jj + 1; c < LUC_W; ++c) x[k][c] -= l * prow[c];
        }
      }
      LUC_TICK(6 + 5 * jj);
    }
  }
  if (swapper && pending) { A[pa] = vb; A[pb2] = va; }
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int li = tid + k * LUC_THREADS;
#pragma unroll
    for (int c = 0; c < LUC_W; ++c)
      if (li < myrows && c < w) a[c * ld + li] = x[k][c];
  }
  __syncthreads();
  LUC_TICK(42);
  for (int e = tid; e < myrows * LUC_W; e += LUC_THREADS) {
    const int li = e >> 3, c = e & 7;
    if (c < w) A[(long long)(j0 + base + li) * lda + j0 + c] = a[c * ld + li];
  }
  LUC_TICK(43);
  if (rank == 0 && sc >= j0 + w && sc < pend) {
    double u[LUC_W];
#pragma unroll
    for (int r = 0; r < LUC_W; ++r) u[r] = r < w ? A[(long long)(j0 + r) * lda + sc] : 0.0;
#pragma unroll
    for (int r = 1; r < LUC_W; ++r) {
      if (r < w) {
#pragma unroll
        for (int q = 0; q < r; ++q) u[r] -= a[q * ld + r] * u[q];
      }
    }
#pragma unroll
    for (int r = 0; r < LUC_W; ++r)
      if (r < w) A[(long long)(j0 + r) * lda + sc] = u[r];
  }
  LUC_TICK(44);
#undef LUC_TICK
}

// A[i][c] -= sum_q A[i][j0+q] A[j0+q][c] for rows i >= j0 + w and the outer panel's columns right of the
// inner panel, c in [j0 + w, pend): a warp per row.
__global__ void __launch_bounds__(256)
lu_inner_update_kernel(double* __restrict__ A, long long lda, int N, int j0, int w, int pend) {
  __shared__ double U[LUC_W][PB];
  const int c0 = j0 + w, ncols = pend - c0;
  for (int t = threadIdx.x; t < LUC_W * ncols; t += blockDim.x) {
    const int q = t / ncols, c = t - q * ncols;
    U[q][c] = q < w ? A[(long long)(j0 + q) * lda + c0 + c] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int i = c0 + blockIdx.x * wpb + warp; i < N; i += gridDim.x * wpb) {
    double* row = A + (long long)i * lda;
    const double mine = lane < w ? row[j0 + lane] : 0.0;
    double l[LUC_W];
#pragma unroll
    for (int q = 0; q < LUC_W; ++q) l[q] = __shfl_sync(0xffffffffu, mine, q);
    for (int c = lane; c < ncols; c += 32) {
      double v = row[c0 + c];
#pragma unroll
      for (int q = 0; q < LUC_W; ++q) v -= l[q] * U[q][c];
      row[c0 + c] = v;
    }
  }
}

// Triangular solve with one diagonal block (pb <= PB rows at k0) of a factorised matrix, unit lower or
// upper, for the columns [c_lo, c_hi) of X (rows k0.. of X; X may be the factorised matrix itself: U12 <-
// inv(L11) A12).  Four lanes share a column (lane = 8 t + column, rows r = 4 i + t in registers), so a warp
// covers eight columns: one round of loads, the substitution fully unrolled in its right-looking form
// (x_q comes from its owner by a shuffle, every lane updates its own rows), one round of stores.  A short
// last block is padded with the identity.
constexpr int LU_SOLVE_THREADS = 256;
constexpr int LU_SOLVE_COLS = LU_SOLVE_THREADS / 4;  // columns per block

template <bool UPPER>
__global__ void __launch_bounds__(LU_SOLVE_THREADS)
lu_diag_solve_reg_kernel(const double* LU, long long lda, int k0, int pb, double* X, long long ldx,
                         int c_lo, int c_hi) {
  static_assert(PB == 64, "row split below assumes a 64-row diagonal block");
  __shared__ double T[PB][PB + 1];
  for (int idx = threadIdx.x; idx < PB * PB; idx += blockDim.x) {
    const int r = idx / PB, c = idx - r * PB;
    double v = r == c ? 1.0 : 0.0;
    if (r < pb && c < pb) v = LU[(long long)(k0 + r) * lda + k0 + c];
    T[r][c] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, cl = lane & 7, t = lane >> 3;
  const int c = c_lo + (blockIdx.x * (LU_SOLVE_THREADS / 32) + (threadIdx.x >> 5)) * 8 + cl;
  const bool live = c < c_hi;  // no early exit: the shuffles below need every lane
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = 4 * i + t;
    x[i] = (live && r < pb) ? X[(long long)(k0 + r) * ldx + c] : 0.0;
  }
  if (!UPPER) {
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const double xq = __shfl_sync(0xffffffffu, x[q >> 2], (q & 3) * 8 + cl);
#pragma unroll
      for (int i = q >> 2; i < 16; ++i)
        if (i > (q >> 2) || t > (q & 3)) x[i] -= T[4 * i + t][q] * xq;
    }
  } else {
#pragma unroll
    for (int q = PB - 1; q >= 0; --q) {
      const double d = x[q >> 2] / T[q][q];
      if (t == (q & 3)) x[q >> 2] = d;
      const double xq = __shfl_sync(0xffffffffu, x[q >> 2], (q & 3) * 8 + cl);
#pragma unroll
      for (int i = 0; i <= (q >> 2); ++i)
        if (i < (q >> 2) || t < (q & 3)) x[i] -= T[4 * i + t][q] * xq;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = 4 * i + t;
    if (live && r < pb) X[(long long)(k0 + r) * ldx + c] = x[i];
  }
}

// sequential row interchanges of the right-hand sides (fallback of getrs for very large N)
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
  // pivot-search partials and the interchange lists (2048 doubles), then the row buffer of the
  // block interchanges: at most 2 LU_OB rows of N entries
  return 2048 + 2LL * LU_OB * std::max(N, 1);
}

static bool g_lu_cluster = true;
extern "C" void pmr_set_lu_cluster_panel(int on) { g_lu_cluster = on != 0; }

namespace {
struct LuScratch {
  double* pval;
  int* pidx;
  int* counter;
  int* count;
  int* list_t;
  int* list_src;
  double* rows;
};

// one PB-wide panel [k0, pend), rows [k0, N): pivots, multipliers, interchanges inside the panel
int lu_factor_panel(double* A, long long lda, int N, int k0, int pend, int* ipiv, int* info,
                    const LuScratch& w, cudaStream_t st) {
  using namespace pmr;
  if (g_lu_cluster && N - k0 <= LUC_CL * LUC_RPC_MAX) {
    // inner panels of LUC_W columns, each one cluster launch (+ one update of the panel's rest)
    for (int j0 = k0; j0 < pend; j0 += LUC_W) {
      const int wd = std::min(LUC_W, pend - j0);
      const int per = (N - j0 + LUC_CL - 1) / LUC_CL;
      const int rpc = (per + 31) / 32 * 32;
      const size_t smem = sizeof(double) * LUC_W * (size_t)(rpc + LUC_PAD);
      if (rpc <= LUC_THREADS)
        lu_inner_panel_kernel<1><<<LUC_CL, LUC_THREADS, smem, st>>>(A, lda, N, k0, pend, j0, wd, rpc,
                                                                    ipiv, info);
      else if (rpc <= 2 * LUC_THREADS)
        lu_inner_panel_kernel<2><<<LUC_CL, LUC_THREADS, smem, st>>>(A, lda, N, k0, pend, j0, wd, rpc,
                                                                    ipiv, info);
      else
        lu_inner_panel_kernel<4><<<LUC_CL, LUC_THREADS, smem, st>>>(A, lda, N, k0, pend, j0, wd, rpc,
                                                                    ipiv, info);
      PMR_CHECK_LAUNCH("lu_inner_panel_kernel");
      if (j0 + wd < pend) {
        const int rows = N - (j0 + wd);
        const int blocks = std::max(1, std::min(LU_MAX_PARTS, (rows + 7) / 8));
        lu_inner_update_kernel<<<blocks, 256, 0, st>>>(A, lda, N, j0, wd, pend);
        PMR_CHECK_LAUNCH("lu_inner_update_kernel");
      }
    }
  } else {
    // one launch per column: finish the previous column and pick the next pivot; one more finishes
    // the panel's last column (nothing to do when the panel reaches the last row)
    for (int j = k0; j <= pend; ++j) {
      const int rows = N - j;
      if (rows <= 0) break;
      const int nparts = std::max(1, std::min(LU_MAX_PARTS, (rows + LU_WARPS - 1) / LU_WARPS));
      lu_column_kernel<<<nparts, LU_WARPS * 32, 0, st>>>(A, lda, N, k0, pend, j, w.pval, w.pidx,
                                                         w.counter, ipiv, info);
      PMR_CHECK_LAUNCH("lu_column_kernel");
    }
  }
  return 0;
}

// interchanges ipiv[p_lo..p_hi) applied to the columns [a_lo, a_hi) and [b_lo, b_hi)
int lu_swap_rows(double* A, long long lda, int N, int p_lo, int p_hi, const int* ipiv, int a_lo,
                 int a_hi, int b_lo, int b_hi, const LuScratch& w, cudaStream_t st) {
  using namespace pmr;
  const int ncols = (a_hi - a_lo) + (b_hi - b_lo);
  if (ncols <= 0 || p_hi <= p_lo) return 0;
  if (N - p_lo > LU_PERM_MAX_ROWS) {
    lu_laswp_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(A, lda, p_lo, p_hi, ipiv, a_lo, a_hi, b_lo,
                                                         b_hi);
    PMR_CHECK_LAUNCH("lu_laswp_kernel");
    return 0;
  }
  lu_perm_build_kernel<<<1, 256, sizeof(int) * (size_t)(N - p_lo), st>>>(ipiv, p_lo, p_hi, N, w.list_t,
                                                                         w.list_src, w.count);
  PMR_CHECK_LAUNCH("lu_perm_build_kernel");
  const dim3 grid((ncols + 255) / 256, 2 * (p_hi - p_lo));
  lu_perm_move_kernel<false><<<grid, 256, 0, st>>>(A, lda, w.list_t, w.list_src, w.count, w.rows, a_lo,
                                                   a_hi, b_lo, b_hi);
  PMR_CHECK_LAUNCH("lu_perm_move_kernel");
  lu_perm_move_kernel<true><<<grid, 256, 0, st>>>(A, lda, w.list_t, w.list_src, w.count, w.rows, a_lo,
                                                  a_hi, b_lo, b_hi);
  PMR_CHECK_LAUNCH("lu_perm_move_kernel");
  return 0;
}

int lu_set_attributes() {
  using namespace pmr;
  static bool done = false;
  if (done) return 0;
  const int luc_smem = (int)(sizeof(double) * LUC_W * (LUC_RPC_MAX + LUC_PAD));
  PMR_CHECK_CUDA(cudaFuncSetAttribute(lu_inner_panel_kernel<1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, luc_smem));
  PMR_CHECK_CUDA(cudaFuncSetAttribute(lu_inner_panel_kernel<2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, luc_smem));
  PMR_CHECK_CUDA(cudaFuncSetAttribute(lu_inner_panel_kernel<4>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, luc_smem));
  PMR_CHECK_CUDA(cudaFuncSetAttribute(lu_perm_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(int) * LU_PERM_MAX_ROWS)));
  PMR_CHECK_CUDA(cudaFuncSetAttribute(lu_perm_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(int) * LU_PERM_MAX_ROWS)));
  done = true;
  return 0;
}
}  // namespace

// Two levels of blocking: PB-wide panels are factorised and applied inside an LU_OB-wide block only;
// the trailing matrix sees one interchange pass, one U12 solve and one rank-LU_OB update per block
// (a quarter of the traffic of rank-PB updates, and GEMMs deep enough to run at the DMMA rate).
extern "C" int pmr_getrf(double* A, int N, long long lda, int* ipiv, int* info, double* work,
                         void* stream) {
  using namespace pmr;
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0, "getrf: bad N");
  PMR_REQUIRE(A && ipiv && info && work, "getrf: null pointer");
  PMR_TRY(lu_set_attributes());
  PMR_CHECK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (N == 0) return 0;
  LuScratch w;
  w.pval = work;
  w.pidx = reinterpret_cast<int*>(work + 1024);
  w.counter = w.pidx + LU_MAX_PARTS;
  w.count = w.pidx + LU_MAX_PARTS + 1;
  w.list_t = w.pidx + 1024;
  w.list_src = w.pidx + 1536;
  w.rows = work + 2048;
  PMR_CHECK_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), st));
  for (int K0 = 0; K0 < N; K0 += LU_OB) {
    const int KEND = std::min(K0 + LU_OB, N);
    for (int k0 = K0; k0 < KEND; k0 += PB) {
      const int pend = std::min(k0 + PB, KEND), pb = pend - k0;
      PMR_TRY(lu_factor_panel(A, lda, N, k0, pend, ipiv, info, w, st));
      PMR_TRY(lu_swap_rows(A, lda, N, k0, pend, ipiv, K0, k0, pend, KEND, w, st));
      if (pend < KEND) {
        const int nc = KEND - pend;
        lu_diag_solve_reg_kernel<false><<<(nc + LU_SOLVE_COLS - 1) / LU_SOLVE_COLS, LU_SOLVE_THREADS, 0, st>>>(A, lda, k0, pb, A, lda, pend,
                                                                          KEND);
        PMR_CHECK_LAUNCH("lu_diag_solve_reg_kernel");
        pmr_gemm_t g = make_gemm(N - pend, nc, pb, -1.0, A + (long long)pend * lda + k0, lda, 1,
                                 A + (long long)k0 * lda + pend, lda, 1, 1.0,
                                 A + (long long)pend * lda + pend, lda, 1);
        PMR_TRY(gemm(g, st));
      }
    }
    PMR_TRY(lu_swap_rows(A, lda, N, K0, KEND, ipiv, 0, K0, KEND, N, w, st));
    const int rest = N - KEND;
    if (rest > 0) {
      // U12 <- inv(L11) A12 with the block's unit-lower L11, PB rows at a time
      for (int k0 = K0; k0 < KEND; k0 += PB) {
        const int pend = std::min(k0 + PB, KEND), pb = pend - k0;
        lu_diag_solve_reg_kernel<false><<<(rest + LU_SOLVE_COLS - 1) / LU_SOLVE_COLS, LU_SOLVE_THREADS, 0, st>>>(A, lda, k0, pb, A, lda, KEND,
                                                                            N);
        PMR_CHECK_LAUNCH("lu_diag_solve_reg_kernel");
        if (pend < KEND) {
          pmr_gemm_t g = make_gemm(KEND - pend, rest, pb, -1.0, A + (long long)pend * lda + k0, lda, 1,
                                   A + (long long)k0 * lda + KEND, lda, 1, 1.0,
                                   A + (long long)pend * lda + KEND, lda, 1);
          PMR_TRY(gemm(g, st));
        }
      }
      pmr_gemm_t g = make_gemm(rest, rest, KEND - K0, -1.0, A + (long long)KEND * lda + K0, lda, 1,
                               A + (long long)K0 * lda + KEND, lda, 1, 1.0,
                               A + (long long)KEND * lda + KEND, lda, 1);
      PMR_TRY(gemm(g, st));
    }
  }
  return 0;
}

extern "C" int pmr_getrs(const double* LU, int N, long long lda, const int* ipiv, double* B,
                         int nrhs, long long ldb, double* work, void* stream) {
  using namespace pmr;
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(N >= 0 && nrhs >= 0, "getrs: bad sizes");
  if (N == 0 || nrhs == 0) return 0;
  PMR_REQUIRE(LU && ipiv && B && work, "getrs: null pointer");
  PMR_TRY(lu_set_attributes());
  if (N <= LU_PERM_MAX_ROWS) {
    // the pivot sequence as one permutation: rows gathered into the work buffer and copied back
    int* perm = reinterpret_cast<int*>(work + (long long)N * nrhs);
    lu_perm_full_kernel<<<1, 1024, sizeof(int) * (size_t)N, st>>>(ipiv, N, perm);
    PMR_CHECK_LAUNCH("lu_perm_full_kernel");
    const dim3 grid((nrhs + 127) / 128, N);
    lu_permute_rows_kernel<<<grid, 128, 0, st>>>(B, ldb, work, nrhs, N, nrhs, perm);
    PMR_CHECK_LAUNCH("lu_permute_rows_kernel");
    lu_permute_rows_kernel<<<grid, 128, 0, st>>>(work, nrhs, B, ldb, N, nrhs, nullptr);
    PMR_CHECK_LAUNCH("lu_permute_rows_kernel");
  } else {
    lu_apply_pivots_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(B, ldb, nrhs, N, ipiv);
    PMR_CHECK_LAUNCH("lu_apply_pivots_kernel");
  }
  // forward (unit lower), right-looking
  for (int k0 = 0; k0 < N; k0 += PB) {
    const int pb = std::min(PB, N - k0), rest = N - k0 - pb;
    lu_diag_solve_reg_kernel<false><<<(nrhs + LU_SOLVE_COLS - 1) / LU_SOLVE_COLS, LU_SOLVE_THREADS, 0, st>>>(LU, lda, k0, pb, B, ldb, 0, nrhs);
    PMR_CHECK_LAUNCH("lu_diag_solve_reg_kernel");
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
    lu_diag_solve_reg_kernel<true><<<(nrhs + LU_SOLVE_COLS - 1) / LU_SOLVE_COLS, LU_SOLVE_THREADS, 0, st>>>(LU, lda, k0, pb, B, ldb, 0, nrhs);
    PMR_CHECK_LAUNCH("lu_diag_solve_reg_kernel");
    if (k0 > 0) {
      pmr_gemm_t g = make_gemm(k0, nrhs, pb, -1.0, LU + k0, lda, 1, B + (long long)k0 * ldb, ldb,
                               1, 1.0, B, ldb, 1);
      PMR_TRY(gemm(g, st));
    }
  }
  return 0;
}
