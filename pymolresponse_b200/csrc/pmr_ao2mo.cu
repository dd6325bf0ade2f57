// AO -> MO two-electron-integral transformation as chains of DMMA GEMMs.
//
// pmr_ao2mo_transform : AO2MO.transform (ao2mo.py:35-50) -- quarters in index order 1,2,3,4.
// pmr_ao2mo_partial   : the occupied-first chain behind perform_rhf_partial (ao2mo.py:58-70)
//                       and perform_uhf_partial (interfaces/pyscf/ao2mo.py:71-109).
//
// Occupied-first ordering (DESIGN.md "K1-K3"): the first quarter (mu -> i, 2 n^4 o flop) is done
// ONCE per bra spin and shared by every block that starts with that spin's occupied index;
// the second contraction is again over an occupied index, so the o*n^3 half-transformed tensor
// is the only large intermediate and it is produced and consumed one nu-block at a time
// (never materialised whole).  No permutational symmetry of I is assumed.
// The AO tensor arrives as a slab I[:, nu_lo:nu_lo+nu_cnt, :, :]: ranks shard the second index of
// the first AO pair, keep mu local (no reduction of the big first-quarter tensor), and sum the
// finished (ia|jb) / (ij|ab) blocks (GBs, not tens of GBs) with one all-reduce each.
#include <algorithm>

#include "pmr_common.cuh"

using namespace pmr;

namespace {

int auto_nu_block(int n, int nu_cnt, int o1, int nu_block) {
  if (nu_block > 0) return std::min(nu_block, nu_cnt);
  // Keep the half-transformed slab H1 = o1 * nb * n^2 doubles within 2^30 doubles (8 GiB).  Every
  // nu-block costs one read-modify-write pass over the accumulators ((ia|jb) and the
  // half-transformed (ij|la si), 15 GB together at n = 512, o = 64), so few large blocks win:
  // with 2 GiB blocks (nb = 16, four passes) the transform ran at 17 TFLOP/s at that size.
  const double per_nu = (double)o1 * n * (double)n;
  int nb = (int)std::max(1.0, std::min((double)nu_cnt, (double)(1LL << 30) / per_nu));
  const long long lim = ((1LL << 31) - 1) / ((long long)n * n);
  nb = (int)std::min<long long>(nb, lim);
  return std::max(1, nb);
}

}  // namespace

extern "C" long long pmr_ao2mo_transform_workspace(int n, int d1, int d2, int d3, int d4) {
  (void)d3; (void)d4;
  const long long n2 = (long long)n * n;
  return (long long)d1 * n2 * n + (long long)d1 * d2 * n2;
}

extern "C" int pmr_ao2mo_transform(const double* I, int n, const double* C1, int ldc1, int d1,
                                   const double* C2, int ldc2, int d2, const double* C3, int ldc3,
                                   int d3, const double* C4, int ldc4, int d4, double* out,
                                   double* workspace, long long workspace_doubles, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 1 && d1 >= 0 && d2 >= 0 && d3 >= 0 && d4 >= 0, "ao2mo_transform: bad dims");
  if (d1 == 0 || d2 == 0 || d3 == 0 || d4 == 0) return 0;
  PMR_REQUIRE(I && C1 && C2 && C3 && C4 && out && workspace, "ao2mo_transform: null pointer");
  PMR_REQUIRE(workspace_doubles >= pmr_ao2mo_transform_workspace(n, d1, d2, d3, d4),
              "ao2mo_transform: workspace too small");
  const long long n2 = (long long)n * n, n3 = n2 * n;
  PMR_REQUIRE(n3 < (1LL << 31) && (long long)d1 * d2 * d3 < (1LL << 31), "ao2mo_transform: too large");
  double* T1 = workspace;                      // (d1, n, n, n), later reused as T3 (d1,d2,d3,n)
  double* T2 = workspace + (long long)d1 * n3; // (d1, d2, n, n)
  // Q1 (ao2mo.py:45): T1[A, nu la si] = sum_mu C1[mu,A] I[mu, nu la si]
  pmr_gemm_t q1 = make_gemm(d1, (int)n3, n, 1.0, C1, 1, ldc1, I, n3, 1, 0.0, T1, n3, 1);
  PMR_TRY(gemm(q1, st));
  // Q2 (ao2mo.py:47): T2[A,B,la si] = sum_nu C2[nu,B] T1[A,nu,la si]
  pmr_gemm_t q2 = make_gemm(d2, (int)n2, n, 1.0, C2, 1, ldc2, T1, n2, 1, 0.0, T2, n2, 1);
  q2.batch1 = d1; q2.a_b1 = 0; q2.b_b1 = n3; q2.c_b1 = (long long)d2 * n2;
  PMR_TRY(gemm(q2, st));
  // Q3 (ao2mo.py:48): T3[A,B,C,si] = sum_la C3[la,C] T2[A,B,la,si]
  double* T3 = T1;
  pmr_gemm_t q3 = make_gemm(d3, n, n, 1.0, C3, 1, ldc3, T2, n, 1, 0.0, T3, n, 1);
  q3.batch1 = d1; q3.batch2 = d2;
  q3.a_b1 = 0; q3.a_b2 = 0;
  q3.b_b1 = (long long)d2 * n2; q3.b_b2 = n2;
  q3.c_b1 = (long long)d2 * d3 * n; q3.c_b2 = (long long)d3 * n;
  PMR_TRY(gemm(q3, st));
  // Q4 (ao2mo.py:49): out[ABC, D] = sum_si T3[ABC, si] C4[si, D]
  pmr_gemm_t q4 = make_gemm((int)((long long)d1 * d2 * d3), d4, n, 1.0, T3, n, 1, C4, ldc4, 1, 0.0,
                            out, d4, 1);
  PMR_TRY(gemm(q4, st));
  return 0;
}

extern "C" long long pmr_ao2mo_partial_workspace(int n, int nu_cnt, int o1, int v1, int n_ket,
                                                 const pmr_ao2mo_ket_t* kets, int want_oovv,
                                                 int nu_block) {
  const int nb = auto_nu_block(n, nu_cnt, o1, nu_block);
  const long long n2 = (long long)n * n;
  long long h1 = (long long)o1 * nb * n2;
  long long x = 0, x2 = 0;
  for (int k = 0; k < n_ket; ++k) {
    x = std::max(x, (long long)o1 * nb * kets[k].o2 * n);
    x2 = std::max(x2, (long long)o1 * nb * kets[k].o2 * kets[k].v2);
  }
  long long y = 0, y2 = 0;
  if (want_oovv) {
    y = (long long)o1 * o1 * n2;
    y2 = (long long)o1 * o1 * n * v1;
  }
  return h1 + x + x2 + y + y2 + 16;
}

extern "C" int pmr_ao2mo_partial(const double* I_slab, int n, int nu_lo, int nu_cnt,
                                 const double* Co1, const double* Cv1, int ldc1, int o1, int v1,
                                 int n_ket, const pmr_ao2mo_ket_t* kets, double* oovv,
                                 int accumulate, int nu_block, double* workspace,
                                 long long workspace_doubles, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 1 && nu_lo >= 0 && nu_cnt >= 0 && nu_lo + nu_cnt <= n, "ao2mo_partial: bad nu range");
  PMR_REQUIRE(o1 >= 0 && v1 >= 0 && n_ket >= 0, "ao2mo_partial: bad dims");
  if (o1 == 0 || nu_cnt == 0) return 0;
  PMR_REQUIRE(I_slab && Co1 && Cv1 && workspace, "ao2mo_partial: null pointer");
  const int want_oovv = oovv != nullptr;
  const bool half_only = (accumulate & 2) != 0;
  PMR_REQUIRE(workspace_doubles >=
                  pmr_ao2mo_partial_workspace(n, nu_cnt, o1, v1, n_ket, kets, want_oovv, nu_block),
              "ao2mo_partial: workspace too small");
  const int nb = auto_nu_block(n, nu_cnt, o1, nu_block);
  const long long n2 = (long long)n * n;
  PMR_REQUIRE((long long)nb * n2 < (1LL << 31), "ao2mo_partial: nu block too large");

  long long xmax = 0, x2max = 0;
  for (int k = 0; k < n_ket; ++k) {
    PMR_REQUIRE(kets[k].Co2 && kets[k].Cv2 && kets[k].ovov, "ao2mo_partial: null ket pointer");
    xmax = std::max(xmax, (long long)o1 * nb * kets[k].o2 * n);
    x2max = std::max(x2max, (long long)o1 * nb * kets[k].o2 * kets[k].v2);
  }
  double* H1 = workspace;
  double* X = H1 + (long long)o1 * nb * n2;
  double* X2 = X + xmax;
  double* Y = X2 + x2max;
  double* Y2 = Y + (want_oovv ? (long long)o1 * o1 * n2 : 0);
  const long long slab_ld = (long long)nu_cnt * n2;  // stride of mu in the slab

  for (int nu0 = 0; nu0 < nu_cnt; nu0 += nb) {
    const int cnt = std::min(nb, nu_cnt - nu0);
    const long long cn2 = (long long)cnt * n2;
    const double beta_blk = ((accumulate & 1) || nu0 > 0) ? 1.0 : 0.0;
    // first quarter, shared: H1[i, nu', la si] = sum_mu Co1[mu,i] I[mu, nu0+nu', la si]
    pmr_gemm_t h = make_gemm(o1, (int)cn2, n, 1.0, Co1, 1, ldc1, I_slab + (long long)nu0 * n2,
                             slab_ld, 1, 0.0, H1, cn2, 1);
    PMR_TRY(gemm(h, st));
    for (int k = 0; k < n_ket; ++k) {
      const pmr_ao2mo_ket_t& kt = kets[k];
      if (kt.o2 == 0 || kt.v2 == 0 || v1 == 0) continue;
      // X[i,nu',j,si] = sum_la Co2[la,j] H1[i,nu',la,si]
      pmr_gemm_t g2 = make_gemm(kt.o2, n, n, 1.0, kt.Co2, 1, kt.ldc2, H1, n, 1, 0.0, X, n, 1);
      g2.batch1 = o1; g2.batch2 = cnt;
      g2.b_b1 = cn2; g2.b_b2 = n2;
      g2.c_b1 = (long long)cnt * kt.o2 * n; g2.c_b2 = (long long)kt.o2 * n;
      PMR_TRY(gemm(g2, st));
      // X2[i,nu',j,b] = sum_si X[i,nu',j,si] Cv2[si,b]
      const long long rows = (long long)o1 * cnt * kt.o2;
      PMR_REQUIRE(rows < (1LL << 31), "ao2mo_partial: too many rows");
      pmr_gemm_t g3 = make_gemm((int)rows, kt.v2, n, 1.0, X, n, 1, kt.Cv2, kt.ldc2, 1, 0.0, X2,
                                kt.v2, 1);
      PMR_TRY(gemm(g3, st));
      // ovov[i,a,j,b] (+)= sum_nu' Cv1[nu_lo+nu0+nu', a] X2[i,nu',j,b]
      const long long jb = (long long)kt.o2 * kt.v2;
      PMR_REQUIRE(jb < (1LL << 31), "ao2mo_partial: o2*v2 too large");
      pmr_gemm_t g4 = make_gemm(v1, (int)jb, cnt, 1.0, Cv1 + (long long)(nu_lo + nu0) * ldc1, 1,
                                ldc1, X2, jb, 1, beta_blk, kt.ovov, jb, 1);
      g4.batch1 = o1; g4.b_b1 = (long long)cnt * jb; g4.c_b1 = (long long)v1 * jb;
      PMR_TRY(gemm(g4, st));
    }
    if (want_oovv) {
      // Y[i,j,la si] (+)= sum_nu' Co1[nu_lo+nu0+nu', j] H1[i,nu',la si]
      pmr_gemm_t y = make_gemm(o1, (int)n2, cnt, 1.0, Co1 + (long long)(nu_lo + nu0) * ldc1, 1, ldc1,
                               H1, n2, 1, nu0 > 0 ? 1.0 : 0.0, Y, n2, 1);
      y.batch1 = o1; y.b_b1 = cn2; y.c_b1 = (long long)o1 * n2;
      PMR_TRY(gemm(y, st));
    }
  }
  if (want_oovv && half_only) {
    // hand the half-transformed (i j|la si) back: the caller reduces it over ranks and finishes
    // only its own rows with pmr_ao2mo_oovv_tail
    PMR_CHECK_CUDA(cudaMemcpyAsync(oovv, Y, sizeof(double) * (size_t)o1 * o1 * n2,
                                   cudaMemcpyDeviceToDevice, st));
  } else if (want_oovv && v1 > 0) {
    PMR_TRY(pmr_ao2mo_oovv_tail(Y, n, Cv1, ldc1, o1, o1, v1, oovv, (accumulate & 1), Y2,
                                (long long)o1 * o1 * n * v1, stream));
  }
  return 0;
}

extern "C" long long pmr_ao2mo_oovv_tail_workspace(int n, int rows, int o1, int v1) {
  return (long long)rows * o1 * n * v1;
}

extern "C" int pmr_ao2mo_oovv_tail(const double* Y, int n, const double* Cv1, int ldc1, int rows,
                                   int o1, int v1, double* oovv, int accumulate, double* workspace,
                                   long long workspace_doubles, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 1 && rows >= 0 && o1 >= 0 && v1 >= 0, "ao2mo_oovv_tail: bad dims");
  if (rows == 0 || o1 == 0 || v1 == 0) return 0;
  PMR_REQUIRE(Y && Cv1 && oovv && workspace, "ao2mo_oovv_tail: null pointer");
  PMR_REQUIRE(workspace_doubles >= pmr_ao2mo_oovv_tail_workspace(n, rows, o1, v1),
              "ao2mo_oovv_tail: workspace too small");
  double* Y2 = workspace;
  // Y2[i,j,la,b] = sum_si Y[i,j,la,si] Cv1[si,b]
  const long long m = (long long)rows * o1 * n;
  PMR_REQUIRE(m < (1LL << 31), "ao2mo_oovv_tail: too many rows");
  pmr_gemm_t g3 = make_gemm((int)m, v1, n, 1.0, Y, n, 1, Cv1, ldc1, 1, 0.0, Y2, v1, 1);
  PMR_TRY(gemm(g3, st));
  // oovv[i,j,a,b] (+)= sum_la Cv1[la,a] Y2[i,j,la,b]
  pmr_gemm_t g4 = make_gemm(v1, v1, n, 1.0, Cv1, 1, ldc1, Y2, v1, 1, accumulate ? 1.0 : 0.0, oovv,
                            v1, 1);
  g4.batch1 = rows; g4.batch2 = o1;
  g4.b_b1 = (long long)o1 * n * v1; g4.b_b2 = (long long)n * v1;
  g4.c_b1 = (long long)o1 * v1 * v1; g4.c_b2 = (long long)v1 * v1;
  PMR_TRY(gemm(g4, st));
  return 0;
}
