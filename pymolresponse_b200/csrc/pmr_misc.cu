// Library state, operator right-hand sides, property contraction, elementwise helpers and the
// one-pass J/K contraction of the matrix-free (iterative) path.
#include <algorithm>
#include <cstring>

#include "pmr_common.cuh"

namespace pmr {

static thread_local char g_error[512] = "";
static thread_local long long g_launches = 0;

char* error_buffer() { return g_error; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches += n; }

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// rhs[c, 0:nov] (already holding scale * vec(C_v^T O_c C_o)) -> mask, mirror with sign, copy
__global__ void rhs_pack_kernel(double* __restrict__ rhs, double* __restrict__ rhs_ai, int ncomp,
                                long long nov, double sign, const int* __restrict__ mask) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= nov || c >= ncomp) return;
  double v = rhs[(long long)c * 2 * nov + k];
  if (mask && mask[k]) v = 0.0;
  rhs[(long long)c * 2 * nov + k] = v;
  rhs[(long long)c * 2 * nov + nov + k] = v * sign;
  if (rhs_ai) rhs_ai[(long long)c * nov + k] = v;
}

constexpr int CONTRACT_CHUNK = 2048;

__global__ void __launch_bounds__(256)
contract_partial_kernel(const double* __restrict__ P, int nP, long long ldp,
                        const double* __restrict__ R, int nR, long long ldr, long long len,
                        double* __restrict__ partial) {
  __shared__ double red[8];
  const long long k0 = (long long)blockIdx.x * CONTRACT_CHUNK;
  const long long k1 = min(k0 + (long long)CONTRACT_CHUNK, len);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a = 0; a < nP; ++a) {
    for (int b = 0; b < nR; ++b) {
      double s = 0.0;
      for (long long k = k0 + threadIdx.x; k < k1; k += 256) s += P[a * ldp + k] * R[b * ldr + k];
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        partial[((long long)blockIdx.x * nP + a) * nR + b] = tot;
      }
      __syncthreads();
    }
  }
}

__global__ void contract_final_kernel(const double* __restrict__ partial, int nblocks, int npairs,
                                      double pref, double* __restrict__ out) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  double s = 0.0;
  for (int q = 0; q < nblocks; ++q) s += partial[(long long)q * npairs + pair];
  out[pair] = pref * s;
}

__global__ void pm_to_xy_kernel(const double* __restrict__ p, const double* __restrict__ m,
                                long long nov, long long ldp, long long ldm,
                                double* __restrict__ xy, long long ldxy) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= nov) return;
  const double pv = p ? p[c * ldp + k] : 0.0;
  const double mv = m ? m[c * ldm + k] : 0.0;
  xy[c * ldxy + k] = 0.5 * (pv + mv);
  xy[c * ldxy + nov + k] = 0.5 * (pv - mv);
}

// ---- packing kernels of the half-size solver (one launch each instead of a dozen tensor ops) ----
constexpr int PACK_MAX_COLS = 64;   // right-hand-side columns (all operators' components)
constexpr int PACK_MAX_FREQ = 64;   // frequencies per launch
struct PackCols { int imag_col[PACK_MAX_COLS]; };   // -1: real operator, else column of T = K^-1 g
struct PackFreq { double w[PACK_MAX_FREQ]; };

// right-hand sides of the m equations, one per ROW of the (rows x N) matrix that is factorised:
//   out[q][c][k] = 2 g[c][k]            (real operator)
//                  2 w_q T[k][imag(c)]  (imaginary operator; T = K^-1 g, (N, nimag))
//                  0                    (padding columns c >= ncols)
__global__ void solve_pack_rhs_kernel(const double* __restrict__ g, long long ldg,
                                      const double* __restrict__ T, int nimag, int ncols, int ncp,
                                      long long N, double* __restrict__ out, long long ldo,
                                      long long stride_q, const PackCols cols, const PackFreq fr) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, q = blockIdx.z;
  if (k >= N) return;
  double v = 0.0;
  if (c < ncols) {
    const int ic = cols.imag_col[c];
    if (ic < 0) v = 2.0 * g[(long long)c * ldg + k];
    else if (T) v = 2.0 * fr.w[q] * T[k * nimag + ic];
  }
  out[(long long)q * stride_q + (long long)c * ldo + k] = v;
  (void)ncp;
}

// right-hand sides of the p equations, all frequencies side by side:
//   prhs[k][q * ncols + c] = w_q m[q][k][c] + (imaginary ? 2 g[c][k] : 0)
__global__ void solve_pack_p_kernel(const double* __restrict__ m, int ncp, const double* __restrict__ g,
                                    long long ldg, int ncols, int nfc, long long N,
                                    double* __restrict__ prhs, const PackCols cols, const PackFreq fr) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_row = (long long)nfc * ncols;
  if (e >= N * per_row) return;
  const long long k = e / per_row;
  const int qc = (int)(e - k * per_row);
  const int q = qc / ncols, c = qc - q * ncols;
  double v = fr.w[q] * m[((long long)q * N + k) * ncp + c];
  if (cols.imag_col[c] >= 0) v += 2.0 * g[(long long)c * ldg + k];
  prhs[e] = v;
}

// xy[(q, c)][0:N] = (p + m) / 2, [N:2N] = (p - m) / 2 with p[k][q * ncols + c] (may be NULL) and
// m[q][k][c]: the (X; Y) supervectors of solvers.py:418-425 for every frequency and column
__global__ void solve_xy_kernel(const double* __restrict__ p, const double* __restrict__ m, int ncp,
                                int ncols, int nfc, long long N, double* __restrict__ xy) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int qc = blockIdx.y;
  if (k >= N) return;
  const int q = qc / ncols, c = qc - q * ncols;
  const double pv = p ? p[k * ((long long)nfc * ncols) + qc] : 0.0;
  const double mv = m[((long long)q * N + k) * ncp + c];
  double* row = xy + (long long)qc * 2 * N;
  row[k] = 0.5 * (pv + mv);
  row[N + k] = 0.5 * (pv - mv);
}

__global__ void axpby_kernel(long long rows, long long cols, double a, const double* __restrict__ x,
                             long long ldx, double b, double* __restrict__ y, long long ldy) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  const double xv = x ? a * x[r * ldx + c] : 0.0;
  y[r * ldy + c] = (b == 0.0) ? xv : xv + b * y[r * ldy + c];
}

__global__ void uncoupled_divide_kernel(const double* __restrict__ sv, long long nov,
                                        const double* __restrict__ ediff, double w,
                                        double* __restrict__ out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= 2 * nov) return;
  const double d = (k < nov) ? (ediff[k] - w) : (ediff[k - nov] + w);
  out[(long long)c * 2 * nov + k] = sv[(long long)c * 2 * nov + k] / d;
}

// One CTA per first AO index p; loops over r, rows q spread over warps, lanes over s.
//   J_d[p,r] = sum_{q,s} I[p,r,q,s] D_d[q,s]      K_d[p,q] = sum_{r,s} I[p,r,q,s] D_d[r,s]
// (J uses (pr|qs) with the roles of the pairs renamed; same tensor element, one pass over I.)
constexpr int JK_MAX_ND = 8;
__global__ void __launch_bounds__(256)
jk_contract_kernel(const double* __restrict__ I, int n, const double* __restrict__ D, int nd,
                   double* __restrict__ J, double* __restrict__ K) {
  extern __shared__ double sm[];  // Ksm[nd][n] + Jred[8][nd]
  double* Ksm = sm;
  double* Jred = sm + (size_t)nd * n;
  const int p = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n2 = (long long)n * n;
  for (int k = threadIdx.x; k < nd * n; k += 256) Ksm[k] = 0.0;
  __syncthreads();
  for (int r = 0; r < n; ++r) {
    const double* Ipr = I + ((long long)p * n + r) * n2;
    double jacc[JK_MAX_ND];
#pragma unroll
    for (int d = 0; d < JK_MAX_ND; ++d) jacc[d] = 0.0;
    for (int q = warp; q < n; q += 8) {
      double kacc[JK_MAX_ND];
#pragma unroll
      for (int d = 0; d < JK_MAX_ND; ++d) kacc[d] = 0.0;
      for (int s = lane; s < n; s += 32) {
        const double v = Ipr[(long long)q * n + s];
#pragma unroll
        for (int d = 0; d < JK_MAX_ND; ++d) {
          if (d < nd) {
            jacc[d] += v * D[d * n2 + (long long)q * n + s];
            kacc[d] += v * D[d * n2 + (long long)r * n + s];
          }
        }
      }
#pragma unroll
      for (int d = 0; d < JK_MAX_ND; ++d) {
        if (d < nd) {
          const double ks = warp_sum(kacc[d]);
          if (lane == 0) Ksm[d * n + q] += ks;  // row q is owned by this warp
        }
      }
    }
#pragma unroll
    for (int d = 0; d < JK_MAX_ND; ++d) {
      if (d < nd) {
        const double js = warp_sum(jacc[d]);
        if (lane == 0) Jred[warp * nd + d] = js;
      }
    }
    __syncthreads();
    if (threadIdx.x < nd) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += Jred[w * nd + threadIdx.x];
      J[threadIdx.x * n2 + (long long)p * n + r] = tot;
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < nd * n; k += 256) {
    const int d = k / n, q = k - d * n;
    K[d * n2 + (long long)p * n + q] = Ksm[k];
  }
}

__global__ void shift_diagonal_kernel(double* __restrict__ A, int N, long long lda, double value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) A[(long long)i * lda + i] += value;
}

// one CTA per component; block-wide max of the signed difference
__global__ void __launch_bounds__(256)
cphf_update_kernel(double* __restrict__ x, const double* __restrict__ x_old,
                   const double* __restrict__ rhs, const double* __restrict__ G,
                   const double* __restrict__ eps, double omega, int n, int nocc, int init,
                   double* __restrict__ maxdiff) {
  __shared__ double red[256];
  const long long n2 = (long long)n * n;
  const long long base = (long long)blockIdx.x * n2;
  double best = -1.0e300;
  for (long long k = threadIdx.x; k < n2; k += 256) {
    const int p = (int)(k / n), q = (int)(k % n);
    double xv = x[base + k];
    const bool ov = (p < nocc) != (q < nocc);
    if (ov && init) {
      // occupied-virtual: rhs / (e_a - e_i - w); virtual-occupied: rhs / -(e_a - e_i - w)
      const double d = (p < nocc) ? (eps[q] - eps[p] - omega) : -(eps[p] - eps[q] - omega);
      xv = rhs[base + k] / d;
      x[base + k] = xv;
    } else if (ov) {
      const double d = eps[p] - eps[q] - omega;
      xv = xv + (rhs[base + k] - (xv * d - G[base + k])) / d;
      x[base + k] = xv;
    }
    best = fmax(best, xv - x_old[base + k]);
  }
  red[threadIdx.x] = best;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) maxdiff[blockIdx.x] = red[0];
}

}  // namespace
}  // namespace pmr

using namespace pmr;

extern "C" int pmr_shift_diagonal(double* A, int N, long long lda, double value, void* stream) {
  PMR_REQUIRE(N >= 0, "shift_diagonal: bad N");
  if (N == 0) return 0;
  PMR_REQUIRE(A, "shift_diagonal: null pointer");
  shift_diagonal_kernel<<<(N + 255) / 256, 256, 0, as_stream(stream)>>>(A, N, lda, value);
  PMR_CHECK_LAUNCH("shift_diagonal_kernel");
  return 0;
}

extern "C" int pmr_cphf_update(double* x, const double* x_old, const double* rhs, const double* G,
                               const double* eps, double omega, int n, int nocc, int ncomp,
                               int init, double* maxdiff, void* stream) {
  PMR_REQUIRE(n >= 1 && nocc >= 0 && nocc <= n && ncomp >= 1, "cphf_update: bad dims");
  PMR_REQUIRE(x && x_old && rhs && G && eps && maxdiff, "cphf_update: null pointer");
  cphf_update_kernel<<<ncomp, 256, 0, as_stream(stream)>>>(x, x_old, rhs, G, eps, omega, n, nocc,
                                                            init, maxdiff);
  PMR_CHECK_LAUNCH("cphf_update_kernel");
  return 0;
}

extern "C" int pmr_version(void) { return 100; }
extern "C" const char* pmr_last_error(void) { return error_buffer(); }
extern "C" long long pmr_launch_count(void) { return g_launches; }
extern "C" void pmr_reset_launch_count(void) { g_launches = 0; }

extern "C" int pmr_form_rhs(const double* O, int ncomp, int n, const double* C, int ldc, int nocc,
                            double sign, double scale, const int* mask, double* rhs, double* rhs_ai,
                            double* workspace, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(ncomp >= 1 && n >= 1 && nocc >= 0 && nocc <= n, "form_rhs: bad dims");
  PMR_REQUIRE(O && C && rhs && workspace, "form_rhs: null pointer");
  const int nv = n - nocc;
  const long long nov = (long long)nocc * nv;
  if (nov == 0) return 0;
  // T_c = O_c C_o   (n x nocc)
  pmr_gemm_t g1 = make_gemm(n, nocc, n, 1.0, O, n, 1, C, ldc, 1, 0.0, workspace, nocc, 1);
  g1.batch1 = ncomp; g1.a_b1 = (long long)n * n; g1.b_b1 = 0; g1.c_b1 = (long long)n * nocc;
  PMR_TRY(gemm(g1, st));
  // G_c = C_v^T T_c (nv x nocc), stored at index i*nv + a  (F-order repack, utils.py:68-70)
  pmr_gemm_t g2 = make_gemm(nv, nocc, n, scale, C + nocc, 1, ldc, workspace, nocc, 1, 0.0, rhs, 1, nv);
  g2.batch1 = ncomp; g2.a_b1 = 0; g2.b_b1 = (long long)n * nocc; g2.c_b1 = 2 * nov;
  PMR_TRY(gemm(g2, st));
  dim3 grid((unsigned)((nov + 255) / 256), ncomp);
  rhs_pack_kernel<<<grid, 256, 0, st>>>(rhs, rhs_ai, ncomp, nov, sign, mask);
  PMR_CHECK_LAUNCH("rhs_pack_kernel");
  return 0;
}

extern "C" long long pmr_contract_workspace(int nP, int nR, long long len) {
  const long long nblocks = (len + CONTRACT_CHUNK - 1) / CONTRACT_CHUNK;
  return std::max(1LL, nblocks) * nP * nR;
}

extern "C" int pmr_contract_results(const double* P, int nP, long long ldp, const double* R, int nR,
                                    long long ldr, long long len, double pref, double* out,
                                    double* workspace, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(nP >= 1 && nR >= 1 && len >= 0, "contract: bad dims");
  PMR_REQUIRE(P && R && out && workspace, "contract: null pointer");
  const int nblocks = (int)std::max(1LL, (len + CONTRACT_CHUNK - 1) / CONTRACT_CHUNK);
  contract_partial_kernel<<<nblocks, 256, 0, st>>>(P, nP, ldp, R, nR, ldr, len, workspace);
  PMR_CHECK_LAUNCH("contract_partial_kernel");
  const int npairs = nP * nR;
  contract_final_kernel<<<(npairs + 127) / 128, 128, 0, st>>>(workspace, nblocks, npairs, pref, out);
  PMR_CHECK_LAUNCH("contract_final_kernel");
  return 0;
}

extern "C" int pmr_pm_to_xy(const double* p, const double* m, int ncomp, long long nov,
                            long long ldp, long long ldm, double* xy, long long ldxy, void* stream) {
  PMR_REQUIRE(ncomp >= 1 && nov >= 0 && xy, "pm_to_xy: bad arguments");
  if (nov == 0) return 0;
  dim3 grid((unsigned)((nov + 255) / 256), ncomp);
  pm_to_xy_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, m, nov, ldp, ldm, xy, ldxy);
  PMR_CHECK_LAUNCH("pm_to_xy_kernel");
  return 0;
}

static int fill_pack(const int* imag_col_host, int ncols, const double* w_host, int nfc, PackCols& pc,
                     PackFreq& pf) {
  PMR_REQUIRE(ncols >= 0 && ncols <= PACK_MAX_COLS, "solve_pack: more than %d right-hand-side columns",
              PACK_MAX_COLS);
  PMR_REQUIRE(nfc >= 0 && nfc <= PACK_MAX_FREQ, "solve_pack: more than %d frequencies per call",
              PACK_MAX_FREQ);
  for (int c = 0; c < PACK_MAX_COLS; ++c) pc.imag_col[c] = (c < ncols && imag_col_host) ? imag_col_host[c] : -1;
  for (int q = 0; q < PACK_MAX_FREQ; ++q) pf.w[q] = (q < nfc && w_host) ? w_host[q] : 0.0;
  return 0;
}

extern "C" int pmr_solve_pack_rhs(const double* g, long long ldg, const double* T, int nimag,
                                  const int* imag_col_host, int ncols, int ncp,
                                  const double* w_host, int nfc, long long N, double* out,
                                  long long ldo, long long stride_q, void* stream) {
  PackCols pc; PackFreq pf;
  PMR_TRY(fill_pack(imag_col_host, ncols, w_host, nfc, pc, pf));
  PMR_REQUIRE(ncp >= ncols && N >= 0 && ldo >= N, "solve_pack_rhs: bad dims");
  if (N == 0 || ncp == 0 || nfc == 0) return 0;
  PMR_REQUIRE(g && out, "solve_pack_rhs: null pointer");
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)ncp, (unsigned)nfc);
  solve_pack_rhs_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, ldg, T, nimag, ncols, ncp, N, out,
                                                             ldo, stride_q, pc, pf);
  PMR_CHECK_LAUNCH("solve_pack_rhs_kernel");
  return 0;
}

extern "C" int pmr_solve_pack_p(const double* m, int ncp, const double* g, long long ldg,
                                const int* imag_col_host, int ncols, const double* w_host, int nfc,
                                long long N, double* prhs, void* stream) {
  PackCols pc; PackFreq pf;
  PMR_TRY(fill_pack(imag_col_host, ncols, w_host, nfc, pc, pf));
  const long long total = N * (long long)nfc * ncols;
  if (total == 0) return 0;
  PMR_REQUIRE(m && g && prhs, "solve_pack_p: null pointer");
  solve_pack_p_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      m, ncp, g, ldg, ncols, nfc, N, prhs, pc, pf);
  PMR_CHECK_LAUNCH("solve_pack_p_kernel");
  return 0;
}

extern "C" int pmr_solve_xy(const double* p, const double* m, int ncp, int ncols, int nfc, long long N,
                            double* xy, void* stream) {
  PMR_REQUIRE(ncols >= 0 && nfc >= 0 && N >= 0 && ncp >= ncols, "solve_xy: bad dims");
  if (N == 0 || ncols == 0 || nfc == 0) return 0;
  PMR_REQUIRE(m && xy, "solve_xy: null pointer");
  PMR_REQUIRE((long long)nfc * ncols <= 65535, "solve_xy: too many columns");
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)(nfc * ncols));
  solve_xy_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, m, ncp, ncols, nfc, N, xy);
  PMR_CHECK_LAUNCH("solve_xy_kernel");
  return 0;
}

extern "C" int pmr_axpby(long long rows, long long cols, double a, const double* x, long long ldx,
                         double b, double* y, long long ldy, void* stream) {
  PMR_REQUIRE(rows >= 0 && cols >= 0 && rows <= 65535, "axpby: bad dims");
  if (rows == 0 || cols == 0) return 0;
  PMR_REQUIRE(y, "axpby: null y");
  dim3 grid((unsigned)((cols + 255) / 256), (unsigned)rows);
  axpby_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, cols, a, x, ldx, b, y, ldy);
  PMR_CHECK_LAUNCH("axpby_kernel");
  return 0;
}

extern "C" int pmr_uncoupled_divide(const double* sv, int ncomp, long long nov, const double* ediff,
                                    double w, double* out, void* stream) {
  PMR_REQUIRE(ncomp >= 1 && nov >= 0, "uncoupled_divide: bad dims");
  if (nov == 0) return 0;
  PMR_REQUIRE(sv && ediff && out, "uncoupled_divide: null pointer");
  dim3 grid((unsigned)((2 * nov + 255) / 256), ncomp);
  uncoupled_divide_kernel<<<grid, 256, 0, as_stream(stream)>>>(sv, nov, ediff, w, out);
  PMR_CHECK_LAUNCH("uncoupled_divide_kernel");
  return 0;
}

extern "C" int pmr_jk_contract(const double* I, int n, const double* D, int nd, double* J, double* K,
                               void* stream) {
  PMR_REQUIRE(n >= 1 && nd >= 1 && nd <= JK_MAX_ND, "jk_contract: need 1 <= nd <= %d", JK_MAX_ND);
  PMR_REQUIRE(I && D && J && K, "jk_contract: null pointer");
  const size_t smem = sizeof(double) * ((size_t)nd * n + 8 * (size_t)nd);
  PMR_REQUIRE(smem <= 48 * 1024, "jk_contract: n too large for the K row buffer");
  jk_contract_kernel<<<n, 256, smem, as_stream(stream)>>>(I, n, D, nd, J, K);
  PMR_CHECK_LAUNCH("jk_contract_kernel");
  return 0;
}
