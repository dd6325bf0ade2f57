// Orbital-Hessian block assembly: one bandwidth-bound gather kernel for all sixteen
// form_rpa_{a,b}_matrix_mo_* functions (explicit_equations_partial.py:8-214,
// explicit_equations_full.py:8-238).
//
//   P[ia,jb] = c1*T1(i,a,j,b) - c2*T2(i,a,j,b) + delta(ia,jb)*(ev[a]-eo[i])
//   Q[ia,jb] = -(c1q*U1(i,a,j,b) - c2q*U2(i,a,j,b))
//
// One CTA produces a 32(a) x 32(b) tile for fixed (i,j).  Each operand is addressed by four
// element strides and read in whichever way is coalesced: directly when b is its unit-stride
// index ((ia|jb) and (ij|ab) both are: the swapaxes(1,2) of explicit_equations_partial.py:35 only
// changes which row is read), or through a padded shared-memory transpose when a is
// ((ib|ja), the swapaxes(1,3) of :93).  CTAs of one (i,j) pair are adjacent in the grid so the
// transposed tile of (ta,tb) is the direct tile of (tb,ta) a few CTAs away: it comes from L2,
// and DRAM traffic stays at the algorithmic 4*nov^2*8 bytes.
// Arithmetic uses __dmul_rn/__dsub_rn/__dadd_rn in the reference's operation order, so the
// blocks are bit-identical to NumPy's (no FMA contraction).
#include <cstdint>
#include <cstdlib>

#include "pmr_common.cuh"

namespace pmr {
namespace {

constexpr int TILE = 32;
constexpr int ROWS_PER_THREAD = 4;  // 256 threads: ty = 0..7 covers rows ty + 8*r

struct Term {
  const double* ptr;
  long long s_i, s_a, s_j, s_b;
  double coef;
  int mode;  // 0 absent, 1 direct (coalesced over b or generic), 2 transposed through smem
};

struct AsmArgs {
  int no_x, nv_x, no_y, nv_y;
  Term t[4];  // p1 p2 q1 q2
  const double* eo;
  const double* ev;
  int out_mode;
  int i_lo, n_i;
  double* out0;
  double* out1;
  long long ld0, ld1;
  int tiles_a, tiles_b;
};

__global__ void __launch_bounds__(256) assemble_kernel(const AsmArgs p) {
  __shared__ double tile[4][TILE][TILE + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  long long bid = blockIdx.x;
  const int tb = (int)(bid % p.tiles_b); bid /= p.tiles_b;
  const int ta = (int)(bid % p.tiles_a); bid /= p.tiles_a;
  const int j = (int)(bid % p.no_y); bid /= p.no_y;
  const int i = p.i_lo + (int)bid;
  const int a0 = ta * TILE, b0 = tb * TILE;

  double val[4][ROWS_PER_THREAD];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const Term& t = p.t[q];
#pragma unroll
    for (int r = 0; r < ROWS_PER_THREAD; ++r) val[q][r] = 0.0;
    if (t.mode == 1) {
      const int b = b0 + tx;
#pragma unroll
      for (int r = 0; r < ROWS_PER_THREAD; ++r) {
        const int a = a0 + ty + 8 * r;
        if (a < p.nv_x && b < p.nv_y)
          val[q][r] = t.ptr[i * t.s_i + a * t.s_a + j * t.s_j + b * t.s_b];
      }
    } else if (t.mode == 2) {
      // a is the unit-stride index: lanes run over a, rows over b, then transpose
      const int a = a0 + tx;
#pragma unroll
      for (int r = 0; r < ROWS_PER_THREAD; ++r) {
        const int b = b0 + ty + 8 * r;
        double v = 0.0;
        if (a < p.nv_x && b < p.nv_y) v = t.ptr[i * t.s_i + a * t.s_a + j * t.s_j + b * t.s_b];
        tile[q][ty + 8 * r][tx] = v;  // [b_local][a_local]
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (p.t[q].mode == 2) {
#pragma unroll
      for (int r = 0; r < ROWS_PER_THREAD; ++r) val[q][r] = tile[q][tx][ty + 8 * r];
    }
  }

  const int b = b0 + tx;
  if (b >= p.nv_y) return;
  const long long jb = (long long)j * p.nv_y + b;
#pragma unroll
  for (int r = 0; r < ROWS_PER_THREAD; ++r) {
    const int a = a0 + ty + 8 * r;
    if (a >= p.nv_x) continue;
    const long long ia = (long long)i * p.nv_x + a;
    double P = __dsub_rn(__dmul_rn(p.t[0].coef, val[0][r]), __dmul_rn(p.t[1].coef, val[1][r]));
    if (p.eo != nullptr && ia == jb) P = __dadd_rn(P, __dsub_rn(p.ev[a], p.eo[i]));
    const double Q =
        -__dsub_rn(__dmul_rn(p.t[2].coef, val[2][r]), __dmul_rn(p.t[3].coef, val[3][r]));
    double o0 = P, o1 = Q;
    if (p.out_mode == 1) {
      o0 = __dsub_rn(P, Q);
      o1 = __dadd_rn(P, Q);
    }
    if (p.out0) p.out0[ia * p.ld0 + jb] = o0;
    if (p.out1) p.out1[ia * p.ld1 + jb] = o1;
  }
}

// Fast path for the partial-layout blocks (everything the solvers assemble at scale): operands are
// either unit-stride in b (read as double2, coalesced 512 B per warp row) or unit-stride in a (read
// coalesced along a, transposed through padded shared memory).  One CTA = a 64(a) x 64(b) tile for
// fixed (i,j); every thread owns 8 rows x one double2 column pair, all (i,j)-dependent address
// arithmetic is hoisted, and the X operand shared by P and Q ((ia|jb)) is loaded once.
// ncu on the generic kernel showed it issue-bound (72 % issue slots, 51 % DRAM): this version
// executes ~2.3x fewer instructions per byte.
constexpr int FT = 64;          // tile edge
constexpr int FLD = FT + 1;     // padded row of the transpose buffer

struct FastArgs {
  int no_x, nv_x, no_y, nv_y;
  const double* X; long long x_si, x_sa, x_sj; double xc_p, xc_q;  // direct, shared by P and Q
  const double* Y; long long y_si, y_sa, y_sj; double yc;          // direct, P only
  const double* Z; long long z_si, z_sb, z_sj; double zc;          // unit stride in a, Q only
  const double* eo; const double* ev;
  int out_mode, i_lo, n_i;
  double* out0; double* out1; long long ld0, ld1;
  int tiles_a, tiles_b;
};

__global__ void __launch_bounds__(256) assemble_fast_kernel(const FastArgs p) {
  __shared__ double tz[FT * FLD];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  long long bid = blockIdx.x;
  const int tb = (int)(bid % p.tiles_b); bid /= p.tiles_b;
  const int ta = (int)(bid % p.tiles_a); bid /= p.tiles_a;
  const int j = (int)(bid % p.no_y); bid /= p.no_y;
  const int i = p.i_lo + (int)bid;
  const int a0 = ta * FT, b0 = tb * FT;
  const int b = b0 + 2 * tx;               // this thread's column pair (nv_y is even)
  const bool bok = b < p.nv_y;

  if (p.Z) {
    // Z(i,a,j,b) = Zbase[i*z_si + b*z_sb + j*z_sj + a]: lanes run over a, rows over b
    const double* zb = p.Z + i * p.z_si + j * p.z_sj;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int al = tx + 32 * h, a = a0 + al;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int bl = ty + 8 * r, bb = b0 + bl;
        double v = 0.0;
        if (a < p.nv_x && bb < p.nv_y) v = zb[bb * p.z_sb + a];
        tz[al * FLD + bl] = v;               // [a_local][b_local]
      }
    }
  }
  double2 x[8], y[8];
  const double* xb = p.X ? p.X + i * p.x_si + j * p.x_sj + b : nullptr;
  const double* yb = p.Y ? p.Y + i * p.y_si + j * p.y_sj + b : nullptr;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int a = a0 + ty + 8 * r;
    x[r] = make_double2(0.0, 0.0);
    y[r] = make_double2(0.0, 0.0);
    if (bok && a < p.nv_x) {
      if (xb) x[r] = *reinterpret_cast<const double2*>(xb + a * p.x_sa);
      if (yb) y[r] = *reinterpret_cast<const double2*>(yb + a * p.y_sa);
    }
  }
  if (p.Z) __syncthreads();
  if (!bok) return;
  const long long jb = (long long)j * p.nv_y + b;
  const double eo_i = p.eo ? p.eo[i] : 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int al = ty + 8 * r, a = a0 + al;
    if (a >= p.nv_x) continue;
    double z0 = 0.0, z1 = 0.0;
    if (p.Z) { z0 = tz[al * FLD + 2 * tx]; z1 = tz[al * FLD + 2 * tx + 1]; }
    const long long ia = (long long)i * p.nv_x + a;
    double P0 = __dsub_rn(__dmul_rn(p.xc_p, x[r].x), __dmul_rn(p.yc, y[r].x));
    double P1 = __dsub_rn(__dmul_rn(p.xc_p, x[r].y), __dmul_rn(p.yc, y[r].y));
    if (p.eo) {
      if (ia == jb) P0 = __dadd_rn(P0, __dsub_rn(p.ev[a], eo_i));
      if (ia == jb + 1) P1 = __dadd_rn(P1, __dsub_rn(p.ev[a], eo_i));
    }
    const double Q0 = -__dsub_rn(__dmul_rn(p.xc_q, x[r].x), __dmul_rn(p.zc, z0));
    const double Q1 = -__dsub_rn(__dmul_rn(p.xc_q, x[r].y), __dmul_rn(p.zc, z1));
    double2 o0 = make_double2(P0, P1), o1 = make_double2(Q0, Q1);
    if (p.out_mode == 1) {
      o0 = make_double2(__dsub_rn(P0, Q0), __dsub_rn(P1, Q1));
      o1 = make_double2(__dadd_rn(P0, Q0), __dadd_rn(P1, Q1));
    }
    if (p.out0) *reinterpret_cast<double2*>(p.out0 + ia * p.ld0 + jb) = o0;
    if (p.out1) *reinterpret_cast<double2*>(p.out1 + ia * p.ld1 + jb) = o1;
  }
}

inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }
inline bool ev2(long long v) { return (v & 1LL) == 0; }

// Can this descriptor run on the fast kernel?  (partial layouts with even nvirt)
bool build_fast(const pmr_assemble_t& a, FastArgs& f) {
  auto direct = [](const pmr_gather_term_t& t) { return t.ptr && t.s_b == 1; };
  auto transposed = [](const pmr_gather_term_t& t) { return t.ptr && t.s_b != 1 && t.s_a == 1; };
  const pmr_gather_term_t &p1 = a.p1, &p2 = a.p2, &q1 = a.q1, &q2 = a.q2;
  if (a.nv_y % 2) return false;
  if (p1.ptr && !direct(p1)) return false;
  if (q1.ptr && !direct(q1)) return false;
  if (p2.ptr && !direct(p2)) return false;
  if (q2.ptr && !transposed(q2)) return false;
  if (p1.ptr && q1.ptr &&
      (p1.ptr != q1.ptr || p1.s_i != q1.s_i || p1.s_a != q1.s_a || p1.s_j != q1.s_j))
    return false;
  const pmr_gather_term_t* x = p1.ptr ? &p1 : (q1.ptr ? &q1 : nullptr);
  f = FastArgs{};
  f.no_x = a.no_x; f.nv_x = a.nv_x; f.no_y = a.no_y; f.nv_y = a.nv_y;
  if (x) {
    if (!al16(x->ptr) || !ev2(x->s_i) || !ev2(x->s_a) || !ev2(x->s_j)) return false;
    f.X = x->ptr; f.x_si = x->s_i; f.x_sa = x->s_a; f.x_sj = x->s_j;
    f.xc_p = p1.ptr ? p1.coef : 0.0;
    f.xc_q = q1.ptr ? q1.coef : 0.0;
  }
  if (p2.ptr) {
    if (!al16(p2.ptr) || !ev2(p2.s_i) || !ev2(p2.s_a) || !ev2(p2.s_j)) return false;
    f.Y = p2.ptr; f.y_si = p2.s_i; f.y_sa = p2.s_a; f.y_sj = p2.s_j; f.yc = p2.coef;
  }
  if (q2.ptr) {
    f.Z = q2.ptr; f.z_si = q2.s_i; f.z_sb = q2.s_b; f.z_sj = q2.s_j; f.zc = q2.coef;
  }
  if (a.out0 && (!al16(a.out0) || !ev2(a.ld0))) return false;
  if (a.out1 && (!al16(a.out1) || !ev2(a.ld1))) return false;
  f.eo = a.eps_occ; f.ev = a.eps_virt;
  f.out_mode = a.out_mode; f.i_lo = a.i_lo; f.n_i = a.i_hi - a.i_lo;
  f.out0 = a.out0; f.out1 = a.out1; f.ld0 = a.ld0; f.ld1 = a.ld1;
  f.tiles_a = (a.nv_x + FT - 1) / FT;
  f.tiles_b = (a.nv_y + FT - 1) / FT;
  return true;
}

__global__ void ediff_kernel(const double* __restrict__ eo, int no, const double* __restrict__ ev,
                             int nv, double* __restrict__ out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (long long)no * nv) return;
  const int i = (int)(k / nv), a = (int)(k % nv);
  out[k] = __dsub_rn(ev[a], eo[i]);
}

Term make_term(const pmr_gather_term_t& g) {
  Term t{};
  t.ptr = g.ptr;
  t.s_i = g.s_i; t.s_a = g.s_a; t.s_j = g.s_j; t.s_b = g.s_b;
  t.coef = g.coef;
  if (!g.ptr) { t.mode = 0; t.coef = 0.0; }
  else if (g.s_b != 1 && g.s_a == 1) t.mode = 2;
  else t.mode = 1;
  return t;
}

}  // namespace
}  // namespace pmr

using namespace pmr;

extern "C" int pmr_assemble(const pmr_assemble_t* a, void* stream) {
  PMR_REQUIRE(a != nullptr, "assemble: null descriptor");
  PMR_REQUIRE(a->no_x >= 0 && a->nv_x >= 0 && a->no_y >= 0 && a->nv_y >= 0, "assemble: bad dims");
  PMR_REQUIRE(a->i_lo >= 0 && a->i_hi <= a->no_x && a->i_lo <= a->i_hi, "assemble: bad row range");
  PMR_REQUIRE(a->out0 || a->out1, "assemble: no output");
  PMR_REQUIRE(a->out_mode == 0 || a->out_mode == 1, "assemble: bad out_mode");
  if (a->eps_occ || a->eps_virt) {
    PMR_REQUIRE(a->eps_occ && a->eps_virt, "assemble: need both eps_occ and eps_virt");
    PMR_REQUIRE(a->no_x == a->no_y && a->nv_x == a->nv_y, "assemble: diagonal on a non-square block");
  }
  static const bool allow_fast = [] { const char* e = getenv("PMR_ASSEMBLE_GENERIC"); return !(e && atoi(e)); }();
  FastArgs f;
  if (allow_fast && build_fast(*a, f)) {
    const long long blocks = (long long)f.n_i * f.no_y * f.tiles_a * f.tiles_b;
    if (blocks == 0) return 0;
    PMR_REQUIRE(blocks < (1LL << 31), "assemble: grid too large");
    assemble_fast_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(f);
    PMR_CHECK_LAUNCH("assemble_fast_kernel");
    return 0;
  }
  AsmArgs p{};
  p.no_x = a->no_x; p.nv_x = a->nv_x; p.no_y = a->no_y; p.nv_y = a->nv_y;
  p.t[0] = make_term(a->p1); p.t[1] = make_term(a->p2);
  p.t[2] = make_term(a->q1); p.t[3] = make_term(a->q2);
  p.eo = a->eps_occ; p.ev = a->eps_virt;
  p.out_mode = a->out_mode;
  p.i_lo = a->i_lo; p.n_i = a->i_hi - a->i_lo;
  p.out0 = a->out0; p.out1 = a->out1; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.tiles_a = (a->nv_x + TILE - 1) / TILE;
  p.tiles_b = (a->nv_y + TILE - 1) / TILE;
  const long long blocks = (long long)p.n_i * p.no_y * p.tiles_a * p.tiles_b;
  if (blocks == 0) return 0;
  PMR_REQUIRE(blocks < (1LL << 31), "assemble: grid too large");
  assemble_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p);
  PMR_CHECK_LAUNCH("assemble_kernel");
  return 0;
}

extern "C" int pmr_energy_differences(const double* eps_occ, int no, const double* eps_virt, int nv,
                                      double* ediff, void* stream) {
  PMR_REQUIRE(no >= 0 && nv >= 0, "energy_differences: bad dims");
  const long long n = (long long)no * nv;
  if (n == 0) return 0;
  PMR_REQUIRE(eps_occ && eps_virt && ediff, "energy_differences: null pointer");
  ediff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(eps_occ, no, eps_virt, nv,
                                                                            ediff);
  PMR_CHECK_LAUNCH("ediff_kernel");
  return 0;
}
