// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM for sm_100a.
//
//   C(m,n) = alpha * sum_k A(m,k) B(k,n) + beta * C(m,n)      (two-level strided batch)
//
// Three kernels share the CTA tile (128x64, 8 math warps of 32x32, BK = 16), the m8n8k4 fragment
// mapping and the epilogue conventions (DESIGN.md section "K-gemm"):
//   * dgemm_tma_kernel  -- both operands k-contiguous (every Cholesky trailing update, C -= P P^T):
//     cp.async.bulk.tensor (tensor-map TMA, 128-byte swizzle, zero-filled edges) issued by one
//     producer lane, full/empty mbarrier ring, no CTA-wide barrier and no global-memory instruction in
//     the math warps' main loop; lower-triangular stores launch only the tiles of the triangle;
//   * dgemm_bulk_kernel -- both operands contiguous along m / n: one 1-D bulk copy per k row;
//   * dgemm_dmma_kernel -- every other layout / alignment: cp.async (LDGSTS) ring with hardware
//     zero-fill, operands in their global layout with +4-double row padding (conflict-free 64-bit
//     fragment loads for both layouts).
// On sm_100a all f64 mma shapes lower to DMMA.8x8x4 (checked with cuobjdump), 16 cycles per SMSP: the
// kernels are DMMA-issue bound by construction (8 LDS.64 per 16 DMMA).  C is written with arbitrary
// (row, column) element strides: transposed / packed-index epilogues (ov packing of
// operators.py:94, quarter-transform output orders) are free.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include <cuda.h>  // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint

#include "pmr_common.cuh"

namespace pmr {

// ---- optional per-launch CUDA-event timing (bench.py roofline pass; off by default) ----
struct ProfRecord { cudaEvent_t e0, e1; int cls; double flops; double bytes; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRecord> g_prof;

namespace {

constexpr int BK = 16;
constexpr int PAD = 4;
// ring depth of the tensor-map kernel: 3 x 24 KiB per CTA, two CTAs per SM (A/B on the config-5
// step: 4 stages 306.4 ms, 3 stages 304.4 ms)
constexpr int TMA_STAGES_64 = 3;

struct GemmArgs {
  int M, N, K;
  double alpha, beta;
  const double* A;
  long long a_ld;  // stride between smem rows of A (m rows if k-contig, k rows otherwise)
  const double* B;
  long long b_ld;
  double* C;
  long long c_sm, c_sn;
  long long a_b1, a_b2, b_b1, b_b2, c_b1, c_b2;
  int tiles_m, tiles_n;
  int flags;
  int a_vec2, b_vec2, c_vec2;
  double* C2;  // optional second (differently strided) copy of the result
  long long c2_sm, c2_sn, c2_b1, c2_b2;
  int tri_rows;  // tensor-map kernel, LOWER: tile rows that are narrower than the matrix
};

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC, int NSTAGE>
struct Cfg {
  static constexpr int WARPS_M = BM / WM;
  static constexpr int WARPS_N = BN / WN;
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8;
  static constexpr int NI = WN / 8;
  static constexpr int LDA = AKC ? (BK + PAD) : (BM + PAD);
  static constexpr int A_ELEMS = AKC ? BM * LDA : BK * LDA;
  static constexpr int LDB = BKC ? (BK + PAD) : (BN + PAD);
  static constexpr int B_ELEMS = BKC ? BN * LDB : BK * LDB;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM = size_t(NSTAGE) * STAGE_ELEMS * sizeof(double);
};

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// Per-thread state for copying one operand tile global -> shared with cp.async.  XT = tile extent
// along the m/n ("x") dimension.  KC: element (x,k) at g[x*ld + k], smem [x][LD]; otherwise element
// at g[k*ld + x], smem [k][LD].  All address arithmetic that does not depend on the k-tile is done
// once here; issuing a stage is then ITER cp.async per thread with one add each, and
// out-of-range elements are zero-filled by the copy engine (src-size < cp-size).
template <int XT, bool KC, int LD, int NT, int VEC>
struct TileLoader {
  static constexpr int CPR = KC ? (BK / VEC) : (XT / VEC);  // chunks per smem row
  static constexpr int TOTAL = KC ? XT * CPR : BK * CPR;
  static constexpr int ITER = (TOTAL + NT - 1) / NT;
  static_assert(NT % CPR == 0, "a thread keeps the same column chunk across its rows");
  static_assert(TOTAL % NT == 0, "every thread copies the same number of chunks");
  static constexpr int ROWS_PER_ITER = NT / CPR;

  const double* src0;   // source of this thread's first chunk in k-tile 0 of the current tile
  int dst0;             // smem offset (elements) of the first chunk
  unsigned mask;        // KC: bit i <=> row of chunk i is inside the matrix; !KC: valid x bytes
  int col;              // KC: k offset of the chunk inside the tile; !KC: first k row of the thread

  __device__ __forceinline__ void init(const double* gbase, long long ld, int x0, int X, int kbase,
                                       int tid) {
    const int r = tid / CPR;
    const int cc = (tid % CPR) * VEC;
    dst0 = r * LD + cc;
    if (KC) {
      col = cc;
      src0 = gbase + (long long)(x0 + r) * ld + kbase + cc;
      mask = 0;
#pragma unroll
      for (int i = 0; i < ITER; ++i)
        if (x0 + r + i * ROWS_PER_ITER < X) mask |= (1u << i);
    } else {
      col = r;
      src0 = gbase + (long long)(kbase + r) * ld + x0 + cc;
      mask = (unsigned)(min(max(X - (x0 + cc), 0), VEC) * 8);
    }
  }

  // kt: k-tile index relative to kbase; k0 = absolute k of the tile; gsafe: any valid address
  __device__ __forceinline__ void issue(double* s, int kt, int k0, int K, long long ld,
                                        const double* gsafe) const {
    const long long step_i = (long long)ROWS_PER_ITER * ld;
    const double* src = src0 + (KC ? (long long)kt * BK : (long long)kt * BK * ld);
    const bool kfull = (k0 + BK <= K);
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
      int bytes;
      if (KC) {
        bytes = ((mask >> i) & 1u) ? (kfull ? VEC * 8 : min(max(K - (k0 + col), 0), VEC) * 8) : 0;
      } else {
        bytes = (kfull || k0 + col + i * ROWS_PER_ITER < K) ? (int)mask : 0;
      }
      const double* ptr = bytes ? (src + i * step_i) : gsafe;
      double* d = s + dst0 + i * ROWS_PER_ITER * LD;
      if (VEC == 2) cp_async16(d, ptr, bytes);
      else cp_async8(d, ptr, bytes);
    }
  }
};

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC, int VEC, int NSTAGE, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) dgemm_dmma_kernel(const GemmArgs p) {
  using C_ = Cfg<BM, BN, WM, WN, AKC, BKC, NSTAGE>;
  constexpr int NT = C_::NT, MI = C_::MI, NI = C_::NI, LDA = C_::LDA, LDB = C_::LDB;
  extern __shared__ __align__(16) double smem[];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tm = blockIdx.x % p.tiles_m, tn = blockIdx.x / p.tiles_m;
  const int bm0 = tm * BM, bn0 = tn * BN;
  const bool lower = (p.flags & PMR_GEMM_LOWER) != 0;
  if (lower && bn0 > bm0 + BM - 1) return;

  const long long b1 = blockIdx.z, b2 = blockIdx.y;
  const double* Ab = p.A + b1 * p.a_b1 + b2 * p.a_b2;
  const double* Bb = p.B + b1 * p.b_b1 + b2 * p.b_b2;
  double* Cb = p.C + b1 * p.c_b1 + b2 * p.c_b2;

  int kt_begin = 0;
  int kt_end = (p.K + BK - 1) / BK;
  if (p.flags & PMR_GEMM_KSTART_M) kt_begin = bm0 / BK;
  if (p.flags & PMR_GEMM_KEND_M) kt_end = min(kt_end, (min(p.K, bm0 + BM) + BK - 1) / BK);
  const int KT = max(kt_end - kt_begin, 0);
  const int kbase = kt_begin * BK;

  const int wm0 = (warp % C_::WARPS_M) * WM;
  const int wn0 = (warp / C_::WARPS_M) * WN;

  // lower-triangular updates: a warp whose whole WM x WN block lies above the diagonal produces
  // nothing that is stored; it keeps copying and synchronising but issues no DMMA, which leaves the
  // tensor pipe to the other CTA of the SM (diagonal tiles are ~2/T of all tiles of a T-row update)
  const bool warp_active = !(lower && bn0 + wn0 > bm0 + wm0 + WM - 1);

  TileLoader<BM, AKC, LDA, NT, VEC> la;  // VEC = 2: 16-byte copies (both operands 16-byte aligned)
  TileLoader<BN, BKC, LDB, NT, VEC> lb;  // VEC = 1: 8-byte copies (odd strides / offsets)
  la.init(Ab, p.a_ld, bm0, p.M, kbase, tid);
  lb.init(Bb, p.b_ld, bn0, p.N, kbase, tid);

  auto load_stage = [&](int slot, int kt) {
    double* As = smem + slot * C_::STAGE_ELEMS;
    double* Bs = As + C_::A_ELEMS;
    const int k0 = kbase + kt * BK;
    la.issue(As, kt, k0, p.K, p.a_ld, p.A);
    lb.issue(Bs, kt, k0, p.K, p.b_ld, p.B);
  };

#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  // beta != 0: the old C tile goes INTO the accumulators (scaled by beta/alpha) while the first
  // stages are in flight, so the epilogue is a pure store: no global-load latency at the tile end.
  const double alpha = p.alpha;
  const bool use_beta = (p.beta != 0.0);
  const double ratio = use_beta ? p.beta / alpha : 0.0;
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int m = bm0 + wm0 + mi * 8 + g;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      if (use_beta) {
        const int n = bn0 + wn0 + ni * 8 + 2 * t;
        if (m < p.M && n < p.N) {
          const bool ok1 = (n + 1 < p.N) && (!lower || n + 1 <= m);
          const bool ok0 = (!lower || n <= m);
          const double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
          if (p.c_vec2 && ok0 && ok1) {
            const double2 o = *reinterpret_cast<const double2*>(c);
            acc[mi][ni][0] = ratio * o.x;
            acc[mi][ni][1] = ratio * o.y;
          } else {
            if (ok0) acc[mi][ni][0] = ratio * c[0];
            if (ok1) acc[mi][ni][1] = ratio * c[p.c_sn];
          }
        }
      }
    }
  }

  int slot = 0;
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    const double* As = smem + slot * C_::STAGE_ELEMS;
    const double* Bs = As + C_::A_ELEMS;
    slot = (slot + 1 == NSTAGE) ? 0 : slot + 1;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      if (warp_active) {
        double a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
          a[mi] = AKC ? As[(wm0 + mi * 8 + g) * LDA + kk * 4 + t]
                      : As[(kk * 4 + t) * LDA + wm0 + mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
          b[ni] = BKC ? Bs[(wn0 + ni * 8 + g) * LDB + kk * 4 + t]
                      : Bs[(kk * 4 + t) * LDB + wn0 + ni * 8 + g];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
      }
      if (kk == 0) {
        // Refill the slot consumed in the previous iteration AFTER the first block of DMMAs is
        // queued: while this warp generates addresses the other warps of the SMSP own the tensor
        // pipe, so copy issue and math overlap instead of alternating in lock step.
        const int nk = kt + NSTAGE - 1;
        if (nk < KT) load_stage((slot + NSTAGE - 2) % NSTAGE, nk);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C(m = ..+g, n = ..+2t, 2t+1) of every 8x8 tile; pure store
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int m = bm0 + wm0 + mi * 8 + g;
    if (m >= p.M) continue;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      const int n = bn0 + wn0 + ni * 8 + 2 * t;
      if (n >= p.N) continue;
      const bool ok1 = (n + 1 < p.N) && (!lower || n + 1 <= m);
      const bool ok0 = (!lower || n <= m);
      double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
      if (p.c_vec2 && ok0 && ok1) {
        double2 v;
        v.x = alpha * acc[mi][ni][0];
        v.y = alpha * acc[mi][ni][1];
        *reinterpret_cast<double2*>(c) = v;
      } else {
        if (ok0) c[0] = alpha * acc[mi][ni][0];
        if (ok1) c[p.c_sn] = alpha * acc[mi][ni][1];
      }
      if (p.C2) {
        double* c2 = p.C2 + b1 * p.c2_b1 + b2 * p.c2_b2 + (long long)m * p.c2_sm + (long long)n * p.c2_sn;
        if (ok0) c2[0] = alpha * acc[mi][ni][0];
        if (ok1) c2[p.c2_sn] = alpha * acc[mi][ni][1];
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC, int VEC, int NSTAGE, int MINB>
int launch_cfg(GemmArgs a, int batch1, int batch2, cudaStream_t st) {
  using C_ = Cfg<BM, BN, WM, WN, AKC, BKC, NSTAGE>;
  auto kern = dgemm_dmma_kernel<BM, BN, WM, WN, AKC, BKC, VEC, NSTAGE, MINB>;
  static thread_local bool configured = false;
  if (!configured) {
    PMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C_::SMEM));
    configured = true;
  }
  a.tiles_m = (a.M + BM - 1) / BM;
  a.tiles_n = (a.N + BN - 1) / BN;
  PMR_REQUIRE((long long)a.tiles_m * a.tiles_n < (1LL << 31), "dgemm: too many tiles");
  PMR_REQUIRE(batch1 <= 65535 && batch2 <= 65535, "dgemm: batch count above 65535");
  dim3 grid((unsigned)((long long)a.tiles_m * a.tiles_n), (unsigned)batch2, (unsigned)batch1);
  ProfRecord rec{};
  if (g_prof_on) {
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    const bool lower = (a.flags & PMR_GEMM_LOWER) != 0;
    const double nb = (double)batch1 * batch2;
    const double nn = std::min(a.M, a.N);
    const double mn = lower ? 0.5 * nn * (nn + 1.0) + ((double)a.M - nn) * nn : (double)a.M * a.N;
    double kk = a.K;
    if (a.flags & PMR_GEMM_KSTART_M) kk = 0.5 * a.K;  // triangular k range (K^-1 = Z^T Z)
    rec.flops = 2.0 * mn * kk * nb;
    rec.bytes = 8.0 * nb * (mn * (a.beta != 0.0 ? 2.0 : 1.0) + (double)a.M * a.K + (double)a.N * a.K);
    rec.cls = (BN == 128 ? 0 : (BN == 64 ? 2 : 4)) + (lower ? 1 : 0);
    cudaEventRecord(rec.e0, st);
  }
  kern<<<grid, C_::NT, C_::SMEM, st>>>(a);
  if (g_prof_on) {
    cudaEventRecord(rec.e1, st);
    g_prof.push_back(rec);
  }
  PMR_CHECK_LAUNCH("dgemm_dmma_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier, warp-specialised variant for operands that are contiguous along
// m / n (A stored [k][m], B stored [k][n]): every k row of a tile is ONE cp.async.bulk (UBLKCP) of up
// to 1 KiB issued by a dedicated producer warp and tracked by the stage's mbarrier (complete_tx);
// the 8/4 math warps never touch global memory or a CTA-wide barrier inside the main loop -- they
// wait on the stage's "full" mbarrier and release the slot through its "empty" mbarrier.
// Used by the first AO->MO quarter (the 34-550 GB AO tensor is the streamed A operand) and by the
// Z^T Z product of the inverse; needs 16-byte aligned rows, even M/N and K % 16 == 0, otherwise the
// cp.async kernel (hardware zero-fill) runs.  Shared-memory layout and fragment mapping are the
// same as above, so are the results (bit-identical accumulation order).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// TMA = true : one producer warp, cp.async.bulk row copies (mn-contiguous operands only pay off);
// TMA = false: two producer warps issue the ordinary 16/8-byte cp.async copies (with zero-fill, any
//              layout / edge) and arrive on the stage's mbarrier with cp.async.mbarrier.arrive.noinc:
//              same decoupled producer/consumer structure for the k-contiguous layouts.
template <int BM, int BN, int WM, int WN, bool AKC, bool BKC, int NSTAGE, int MINB, bool TMA, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32 + (TMA ? 32 : 64), MINB)
dgemm_bulk_kernel(const GemmArgs p) {
  using C_ = Cfg<BM, BN, WM, WN, AKC, BKC, NSTAGE>;
  constexpr int NTC = C_::NT;  // consumer threads
  constexpr int NWC = NTC / 32;
  constexpr int MI = C_::MI, NI = C_::NI, LDA = C_::LDA, LDB = C_::LDB;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long full_bar[NSTAGE];
  __shared__ __align__(8) unsigned long long empty_bar[NSTAGE];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tm = blockIdx.x % p.tiles_m, tn = blockIdx.x / p.tiles_m;
  const int bm0 = tm * BM, bn0 = tn * BN;
  const bool lower = (p.flags & PMR_GEMM_LOWER) != 0;
  if (lower && bn0 > bm0 + BM - 1) return;

  const long long b1 = blockIdx.z, b2 = blockIdx.y;
  const double* Ab = p.A + b1 * p.a_b1 + b2 * p.a_b2;
  const double* Bb = p.B + b1 * p.b_b1 + b2 * p.b_b2;
  double* Cb = p.C + b1 * p.c_b1 + b2 * p.c_b2;

  constexpr int NPROD = TMA ? 32 : 64;  // producer threads
  int kt_begin = 0;
  int kt_end = (p.K + BK - 1) / BK;  // (K % BK == 0 on the TMA path)
  if (p.flags & PMR_GEMM_KSTART_M) kt_begin = bm0 / BK;
  if (p.flags & PMR_GEMM_KEND_M) kt_end = min(kt_end, (min(p.K, bm0 + BM) + BK - 1) / BK);
  const int KT = max(kt_end - kt_begin, 0);
  const int kbase = kt_begin * BK;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], TMA ? 1 : NPROD);
      mbar_init(&empty_bar[s], NWC);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if constexpr (!TMA) {
  if (warp >= NWC) {
    // ===== producer warps, cp.async flavour =====
    const int ptid = tid - NTC;
    TileLoader<BM, AKC, LDA, NPROD, VEC> la;
    TileLoader<BN, BKC, LDB, NPROD, VEC> lb;
    la.init(Ab, p.a_ld, bm0, p.M, kbase, ptid);
    lb.init(Bb, p.b_ld, bn0, p.N, kbase, ptid);
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % NSTAGE;
      if (kt >= NSTAGE) mbar_wait(&empty_bar[slot], ((kt / NSTAGE) - 1) & 1);
      double* As = smem + slot * C_::STAGE_ELEMS;
      double* Bs = As + C_::A_ELEMS;
      const int k0 = kbase + kt * BK;
      la.issue(As, kt, k0, p.K, p.a_ld, p.A);
      lb.issue(Bs, kt, k0, p.K, p.b_ld, p.B);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full_bar[slot]))
                   : "memory");
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    return;
  }
  }
  if (TMA && warp == NWC) {
    // ===== producer warp, TMA bulk flavour =====
    // mn-contiguous operand: one copy per k row (16 rows of up to BM*8 bytes);
    // k-contiguous operand : one 128-byte copy per m/n row of the tile.
    const int rows_a = min(BM, p.M - bm0), rows_b = min(BN, p.N - bn0);
    const unsigned tx_a = AKC ? (unsigned)rows_a * (BK * 8) : (unsigned)BK * (rows_a * 8);
    const unsigned tx_b = BKC ? (unsigned)rows_b * (BK * 8) : (unsigned)BK * (rows_b * 8);
    const double* a_src = AKC ? Ab + (long long)bm0 * p.a_ld + kbase : Ab + (long long)kbase * p.a_ld + bm0;
    const double* b_src = BKC ? Bb + (long long)bn0 * p.b_ld + kbase : Bb + (long long)kbase * p.b_ld + bn0;
    const long long a_step = AKC ? (long long)BK : (long long)BK * p.a_ld;
    const long long b_step = BKC ? (long long)BK : (long long)BK * p.b_ld;
    for (int kt = 0; kt < KT; ++kt) {
      const int slot = kt % NSTAGE;
      if (kt >= NSTAGE) mbar_wait(&empty_bar[slot], ((kt / NSTAGE) - 1) & 1);
      if (lane == 0) mbar_expect_tx(&full_bar[slot], tx_a + tx_b);
      __syncwarp();
      double* As = smem + slot * C_::STAGE_ELEMS;
      double* Bs = As + C_::A_ELEMS;
      const double* ak = a_src + kt * a_step;
      const double* bk = b_src + kt * b_step;
      if (AKC) {
        for (int r = lane; r < rows_a; r += 32)
          bulk_g2s(As + r * LDA, ak + (long long)r * p.a_ld, BK * 8, &full_bar[slot]);
      } else if (lane < BK) {
        bulk_g2s(As + lane * LDA, ak + (long long)lane * p.a_ld, rows_a * 8, &full_bar[slot]);
      }
      if (BKC) {
        for (int r = lane; r < rows_b; r += 32)
          bulk_g2s(Bs + r * LDB, bk + (long long)r * p.b_ld, BK * 8, &full_bar[slot]);
      } else if (lane >= 32 - BK) {
        const int kr = lane - (32 - BK);
        bulk_g2s(Bs + kr * LDB, bk + (long long)kr * p.b_ld, rows_b * 8, &full_bar[slot]);
      }
    }
    return;
  }

  // ===== math warps =====
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % C_::WARPS_M) * WM;
  const int wn0 = (warp / C_::WARPS_M) * WN;
  const double alpha = p.alpha;
  const bool use_beta = (p.beta != 0.0);
  const double ratio = use_beta ? p.beta / alpha : 0.0;
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int m = bm0 + wm0 + mi * 8 + g;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      if (use_beta) {
        const int n = bn0 + wn0 + ni * 8 + 2 * t;
        if (m < p.M && n < p.N) {
          const bool ok1 = (n + 1 < p.N) && (!lower || n + 1 <= m);
          const bool ok0 = (!lower || n <= m);
          const double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
          if (p.c_vec2 && ok0 && ok1) {
            const double2 o = *reinterpret_cast<const double2*>(c);
            acc[mi][ni][0] = ratio * o.x;
            acc[mi][ni][1] = ratio * o.y;
          } else {
            if (ok0) acc[mi][ni][0] = ratio * c[0];
            if (ok1) acc[mi][ni][1] = ratio * c[p.c_sn];
          }
        }
      }
    }
  }

  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % NSTAGE;
    mbar_wait(&full_bar[slot], (kt / NSTAGE) & 1);
    const double* As = smem + slot * C_::STAGE_ELEMS;
    const double* Bs = As + C_::A_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[MI], b[NI];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
        a[mi] = AKC ? As[(wm0 + mi * 8 + g) * LDA + kk * 4 + t]
                    : As[(kk * 4 + t) * LDA + wm0 + mi * 8 + g];
#pragma unroll
      for (int ni = 0; ni < NI; ++ni)
        b[ni] = BKC ? Bs[(wn0 + ni * 8 + g) * LDB + kk * 4 + t]
                    : Bs[(kk * 4 + t) * LDB + wn0 + ni * 8 + g];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp is done reading the slot
  }

#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int m = bm0 + wm0 + mi * 8 + g;
    if (m >= p.M) continue;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      const int n = bn0 + wn0 + ni * 8 + 2 * t;
      if (n >= p.N) continue;
      const bool ok1 = (n + 1 < p.N) && (!lower || n + 1 <= m);
      const bool ok0 = (!lower || n <= m);
      double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
      if (p.c_vec2 && ok0 && ok1) {
        double2 v;
        v.x = alpha * acc[mi][ni][0];
        v.y = alpha * acc[mi][ni][1];
        *reinterpret_cast<double2*>(c) = v;
      } else {
        if (ok0) c[0] = alpha * acc[mi][ni][0];
        if (ok1) c[p.c_sn] = alpha * acc[mi][ni][1];
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC, int NSTAGE, int MINB, bool TMA, int VEC>
int launch_bulk_cfg(GemmArgs a, int batch1, int batch2, cudaStream_t st) {
  using C_ = Cfg<BM, BN, WM, WN, AKC, BKC, NSTAGE>;
  auto kern = dgemm_bulk_kernel<BM, BN, WM, WN, AKC, BKC, NSTAGE, MINB, TMA, VEC>;
  static thread_local bool configured = false;
  if (!configured) {
    PMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C_::SMEM));
    configured = true;
  }
  a.tiles_m = (a.M + BM - 1) / BM;
  a.tiles_n = (a.N + BN - 1) / BN;
  PMR_REQUIRE((long long)a.tiles_m * a.tiles_n < (1LL << 31), "dgemm: too many tiles");
  PMR_REQUIRE(batch1 <= 65535 && batch2 <= 65535, "dgemm: batch count above 65535");
  dim3 grid((unsigned)((long long)a.tiles_m * a.tiles_n), (unsigned)batch2, (unsigned)batch1);
  ProfRecord rec{};
  if (g_prof_on) {
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    const bool lower = (a.flags & PMR_GEMM_LOWER) != 0;
    const double nb = (double)batch1 * batch2;
    const double nn = std::min(a.M, a.N);
    const double mn = lower ? 0.5 * nn * (nn + 1.0) + ((double)a.M - nn) * nn : (double)a.M * a.N;
    double kk = a.K;
    if (a.flags & PMR_GEMM_KSTART_M) kk = 0.5 * a.K;
    rec.flops = 2.0 * mn * kk * nb;
    rec.bytes = 8.0 * nb * (mn * (a.beta != 0.0 ? 2.0 : 1.0) + (double)a.M * a.K + (double)a.N * a.K);
    rec.cls = (BN == 128 ? 0 : (BN == 64 ? 2 : 4)) + (lower ? 1 : 0);
    cudaEventRecord(rec.e0, st);
  }
  kern<<<grid, C_::NT + (TMA ? 32 : 64), C_::SMEM, st>>>(a);
  if (g_prof_on) {
    cudaEventRecord(rec.e1, st);
    g_prof.push_back(rec);
  }
  PMR_CHECK_LAUNCH("dgemm_bulk_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// TMA tensor-map variant for k-CONTIGUOUS operands (A stored [m][k], B stored [n][k]): the layout of
// every Cholesky trailing update (C -= P P^T), the dominant kernel of the step.  One elected lane of a
// producer warp issues ONE cp.async.bulk.tensor (SASS UTMALDG) per operand and stage -- a BM x 16 and
// a BN x 16 box of a 4-D tensor map {k, row, batch2, batch1} -- tracked by the stage's "full"
// mbarrier; out-of-range rows / k are zero-filled by the TMA unit, so every M/N/K edge is legal.
// Shared memory is dense (128-byte rows, no padding) with the hardware 128-byte swizzle: the 16-byte
// chunk c of row r lives at chunk c ^ (r & 7), which makes the per-lane 64-bit fragment loads of the
// m8n8k4 shapes bank-conflict free (8 rows x 32 bytes of a warp's request fall on 4 distinct chunk
// pairs, two rows each: the two-wavefront minimum of a 256-byte request; LDS.64 is served per half
// warp, so ncu still counts 3.1 wavefronts per load.  A row permutation {0,2,4,6,1,3,5,7} of the
// fragment slots removes them (757 M -> 27 M conflicts) and gains 1.8 % on full beta = 0 products, but
// nothing on the lower-triangular beta = 1 updates this kernel exists for, whose epilogue it
// complicates (-2.5 %): measured and not kept).  The math warps never
// touch global memory or a CTA-wide barrier in the main loop, exactly as in dgemm_bulk_kernel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%2, %3, %4, %5}], [%6];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(smem_u32(bar))
      : "memory");
}

template <int BM, int BN, int WM, int WN, int NSTAGE, bool LOWER>
__device__ __forceinline__ void dgemm_tma_body(const CUtensorMap& mapA, const CUtensorMap& mapB,
                                               const GemmArgs& p) {
  constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  constexpr int NWC = WARPS_M * WARPS_N;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int ROWB = BK * 8;                       // 128 bytes per tile row
  constexpr int A_BYTES = BM * ROWB, B_BYTES = BN * ROWB, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[NSTAGE];
  __shared__ __align__(8) unsigned long long empty_bar[NSTAGE];
  // the swizzle pattern is a function of the shared-memory ADDRESS: tiles start on 1 KiB boundaries
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr bool lower = LOWER;
  int tm, tn;
  if (lower) {
    // only the tiles that touch the lower triangle are launched: row tm has R (tm + 1) of them
    // (R = BM / BN) until the row is as wide as the matrix; tri_rows rows are of that kind
    constexpr int R = (BM >= BN) ? BM / BN : 1;
    const int bid = (int)blockIdx.x;
    const int tri_count = R * p.tri_rows * (p.tri_rows + 1) / 2;
    if (bid < tri_count) {
      tm = (int)((sqrtf(1.0f + 8.0f * (float)bid / (float)R) - 1.0f) * 0.5f);
      while (R * (tm + 1) * (tm + 2) / 2 <= bid) ++tm;
      while (R * tm * (tm + 1) / 2 > bid) --tm;
      tn = bid - R * tm * (tm + 1) / 2;
    } else {
      const int rest = bid - tri_count;
      tm = p.tri_rows + rest / p.tiles_n;
      tn = rest % p.tiles_n;
    }
  } else {
    tm = blockIdx.x % p.tiles_m;
    tn = blockIdx.x / p.tiles_m;
  }
  const int bm0 = tm * BM, bn0 = tn * BN;

  const long long b1 = blockIdx.z, b2 = blockIdx.y;
  double* Cb = p.C + b1 * p.c_b1 + b2 * p.c_b2;

  int kt_begin = 0;
  int kt_end = (p.K + BK - 1) / BK;
  if (p.flags & PMR_GEMM_KSTART_M) kt_begin = bm0 / BK;
  if (p.flags & PMR_GEMM_KEND_M) kt_end = min(kt_end, (min(p.K, bm0 + BM) + BK - 1) / BK);
  const int KT = max(kt_end - kt_begin, 0);
  const int kbase = kt_begin * BK;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NWC);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == NWC) {
    // ===== producer warp: one elected lane feeds the ring =====
    if (lane == 0) {
      // a batch stride of 0 (one operand shared by the whole batch) is a dimension of extent 1
      const int a2 = p.a_b2 ? (int)b2 : 0, a1 = p.a_b1 ? (int)b1 : 0;
      const int bb2 = p.b_b2 ? (int)b2 : 0, bb1 = p.b_b1 ? (int)b1 : 0;
      for (int kt = 0; kt < KT; ++kt) {
        const int slot = kt % NSTAGE;
        if (kt >= NSTAGE) mbar_wait(&empty_bar[slot], ((kt / NSTAGE) - 1) & 1);
        mbar_expect_tx(&full_bar[slot], STAGE_BYTES);
        unsigned char* As = tiles + slot * STAGE_BYTES;
        const int k0 = kbase + kt * BK;
        tma_load_4d(As, &mapA, k0, bm0, a2, a1, &full_bar[slot]);
        tma_load_4d(As + A_BYTES, &mapB, k0, bn0, bb2, bb1, &full_bar[slot]);
      }
    }
    return;
  }

  // ===== math warps =====
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % WARPS_M) * WM;
  const int wn0 = (warp / WARPS_M) * WN;
  const bool warp_active = !(lower && bn0 + wn0 > bm0 + wm0 + WM - 1);
  const double alpha = p.alpha;
  const bool use_beta = (p.beta != 0.0);
  // beta != 0: the old C tile is only PREFETCHED into L2 here and added in the epilogue, so the first
  // DMMA never waits for a DRAM round trip (measured on the config-5 step: 321.3 -> 311.3 ms against
  // loading it into the accumulators up front)
  if (use_beta && p.c_vec2 && warp_active) {
#pragma unroll
    for (int h = 0; h < WM / 16; ++h) {
      const int m = bm0 + wm0 + (lane >> 1) + 16 * h;
      const int n = bn0 + wn0 + (lane & 1) * 16;
      if (m < p.M && n < p.N && (!lower || n <= m))
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Cb + (long long)m * p.c_sm + n));
    }
  }
  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  // byte offset of element (row = 8 q + g, k = 4 kk + t) inside a tile: rows are 128 bytes, the
  // 16-byte chunk (2 kk + t/2) is XOR-ed with g (= row & 7: wm0, wn0 and 8 q are multiples of 8)
  // (the k-step only flips address bits 5-6: offset(kk) = offset(0) ^ (kk << 5), one register)
  const int koff0 = g * ROWB + (((t >> 1) ^ g) << 4) + ((t & 1) << 3);

  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % NSTAGE;
    mbar_wait(&full_bar[slot], (kt / NSTAGE) & 1);
    if (warp_active) {
      const unsigned char* As = tiles + slot * STAGE_BYTES + wm0 * ROWB;
      const unsigned char* Bs = tiles + slot * STAGE_BYTES + A_BYTES + wn0 * ROWB;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MI], b[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
          a[mi] = *reinterpret_cast<const double*>(As + mi * 8 * ROWB + (koff0 ^ (kk << 5)));
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
          b[ni] = *reinterpret_cast<const double*>(Bs + ni * 8 * ROWB + (koff0 ^ (kk << 5)));
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp is done reading the slot
  }
  if (!warp_active) return;

  const double ratio = use_beta ? p.beta / alpha : 0.0;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int m = bm0 + wm0 + mi * 8 + g;
    if (m >= p.M) continue;
    double2 old[NI];
    if (use_beta) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int n = bn0 + wn0 + ni * 8 + 2 * t;
        old[ni] = make_double2(0.0, 0.0);
        if (n < p.N && (!lower || n <= m)) {
          const double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
          const bool two = n + 1 < p.N && (!lower || n + 1 <= m);
          if (p.c_vec2 && two) old[ni] = *reinterpret_cast<const double2*>(c);
          else {
            old[ni].x = c[0];
            if (two) old[ni].y = c[p.c_sn];
          }
        }
      }
    }
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      const int n = bn0 + wn0 + ni * 8 + 2 * t;
      if (n >= p.N) continue;
      const bool ok1 = (n + 1 < p.N) && (!lower || n + 1 <= m);
      const bool ok0 = (!lower || n <= m);
      double* c = Cb + (long long)m * p.c_sm + (long long)n * p.c_sn;
      double2 v;
      v.x = acc[mi][ni][0];
      v.y = acc[mi][ni][1];
      if (use_beta) { v.x = fma(ratio, old[ni].x, v.x); v.y = fma(ratio, old[ni].y, v.y); }
      v.x *= alpha;
      v.y *= alpha;
      if (p.c_vec2 && ok0 && ok1) {
        *reinterpret_cast<double2*>(c) = v;
      } else {
        if (ok0) c[0] = v.x;
        if (ok1) c[p.c_sn] = v.y;
      }
    }
  }
}

// (ptxas caps the 288-thread kernel at 96 registers for two CTAs per SM, with a few spills in the k
// loop; a __maxnreg__(112) build fits the register file on paper but ran 9 % slower: not kept.)
template <int BM, int BN, int WM, int WN, int NSTAGE, int MINB, bool LOWER>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32 + 32, MINB)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const GemmArgs p) {
  dgemm_tma_body<BM, BN, WM, WN, NSTAGE, LOWER>(mapA, mapB, p);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

// 4-D map {k, rows, batch2, batch1} over a k-contiguous operand; box = 16 x box_rows x 1 x 1
static bool make_operand_map(CUtensorMap* map, const double* base, int rows, int K, long long ld,
                             int batch2, long long s_b2, int batch1, long long s_b1, int box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t fallback = (cuuint64_t)ld * 8ull * (cuuint64_t)std::max(rows, 1);
  cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(s_b2 ? batch2 : 1),
                        (cuuint64_t)(s_b1 ? batch1 : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 8ull, s_b2 ? (cuuint64_t)s_b2 * 8ull : fallback,
                           s_b1 ? (cuuint64_t)s_b1 * 8ull : fallback};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0 || strides[i] >= (1ull << 40)) return false;
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// returns 0 on launch, >0 when the operands cannot be described by a tensor map (caller falls back
// to the cp.async kernel), <0 on error
template <int BM, int BN, int WM, int WN, int NSTAGE, int MINB, bool LOWER>
int launch_tma_lower(GemmArgs a, int batch1, int batch2, cudaStream_t st) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32 + 32;
  constexpr int SMEM = NSTAGE * (BM + BN) * BK * 8 + 1024;
  CUtensorMap mapA, mapB;
  if (!make_operand_map(&mapA, a.A, a.M, a.K, a.a_ld, batch2, a.a_b2, batch1, a.a_b1, BM) ||
      !make_operand_map(&mapB, a.B, a.N, a.K, a.b_ld, batch2, a.b_b2, batch1, a.b_b1, BN))
    return 1;
  auto kern = dgemm_tma_kernel<BM, BN, WM, WN, NSTAGE, MINB, LOWER>;
  static thread_local bool configured = false;
  if (!configured) {
    PMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  a.tiles_m = (a.M + BM - 1) / BM;
  a.tiles_n = (a.N + BN - 1) / BN;
  PMR_REQUIRE((long long)a.tiles_m * a.tiles_n < (1LL << 31), "dgemm: too many tiles");
  PMR_REQUIRE(batch1 <= 65535 && batch2 <= 65535, "dgemm: batch count above 65535");
  long long ntiles = (long long)a.tiles_m * a.tiles_n;
  if (a.flags & PMR_GEMM_LOWER) {
    constexpr int R = (BM >= BN) ? BM / BN : 1;
    a.tri_rows = std::min(a.tiles_m, a.tiles_n / R);     // rows with R (tm + 1) <= tiles_n
    ntiles = (long long)R * a.tri_rows * (a.tri_rows + 1) / 2 + (long long)(a.tiles_m - a.tri_rows) * a.tiles_n;
  }
  dim3 grid((unsigned)ntiles, (unsigned)batch2, (unsigned)batch1);
  ProfRecord rec{};
  if (g_prof_on) {
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    const bool lower = (a.flags & PMR_GEMM_LOWER) != 0;
    const double nb = (double)batch1 * batch2;
    const double nn = std::min(a.M, a.N);
    const double mn = lower ? 0.5 * nn * (nn + 1.0) + ((double)a.M - nn) * nn : (double)a.M * a.N;
    double kk = a.K;
    if (a.flags & PMR_GEMM_KSTART_M) kk = 0.5 * a.K;
    rec.flops = 2.0 * mn * kk * nb;
    rec.bytes = 8.0 * nb * (mn * (a.beta != 0.0 ? 2.0 : 1.0) + (double)a.M * a.K + (double)a.N * a.K);
    rec.cls = (BN == 128 ? 0 : (BN == 64 ? 2 : 4)) + (lower ? 1 : 0);
    cudaEventRecord(rec.e0, st);
  }
  kern<<<grid, NT, SMEM, st>>>(mapA, mapB, a);
  if (g_prof_on) {
    cudaEventRecord(rec.e1, st);
    g_prof.push_back(rec);
  }
  PMR_CHECK_LAUNCH("dgemm_tma_kernel");
  return 0;
}

template <int BM, int BN, int WM, int WN, int NSTAGE, int MINB>
int launch_tma_cfg(const GemmArgs& a, int batch1, int batch2, cudaStream_t st) {
  if (a.flags & PMR_GEMM_LOWER) return launch_tma_lower<BM, BN, WM, WN, NSTAGE, MINB, true>(a, batch1, batch2, st);
  return launch_tma_lower<BM, BN, WM, WN, NSTAGE, MINB, false>(a, batch1, batch2, st);
}

template <int BM, int BN, int WM, int WN, int NSTAGE, int MINB>
int launch_bulk(const GemmArgs& a, bool akc, bool bkc, int b1, int b2, cudaStream_t st) {
  if (akc && bkc) return launch_bulk_cfg<BM, BN, WM, WN, true, true, NSTAGE, MINB, true, 2>(a, b1, b2, st);
  if (akc && !bkc) return launch_bulk_cfg<BM, BN, WM, WN, true, false, NSTAGE, MINB, true, 2>(a, b1, b2, st);
  if (!akc && bkc) return launch_bulk_cfg<BM, BN, WM, WN, false, true, NSTAGE, MINB, true, 2>(a, b1, b2, st);
  return launch_bulk_cfg<BM, BN, WM, WN, false, false, NSTAGE, MINB, true, 2>(a, b1, b2, st);
}

template <int BM, int BN, int WM, int WN, int NSTAGE, int MINB>
int launch_layout(const GemmArgs& a, bool akc, bool bkc, int b1, int b2, cudaStream_t st) {
  if (a.a_vec2 && a.b_vec2) {
    if (akc && bkc) return launch_cfg<BM, BN, WM, WN, true, true, 2, NSTAGE, MINB>(a, b1, b2, st);
    if (akc && !bkc) return launch_cfg<BM, BN, WM, WN, true, false, 2, NSTAGE, MINB>(a, b1, b2, st);
    if (!akc && bkc) return launch_cfg<BM, BN, WM, WN, false, true, 2, NSTAGE, MINB>(a, b1, b2, st);
    return launch_cfg<BM, BN, WM, WN, false, false, 2, NSTAGE, MINB>(a, b1, b2, st);
  }
  if (akc && bkc) return launch_cfg<BM, BN, WM, WN, true, true, 1, NSTAGE, MINB>(a, b1, b2, st);
  if (akc && !bkc) return launch_cfg<BM, BN, WM, WN, true, false, 1, NSTAGE, MINB>(a, b1, b2, st);
  if (!akc && bkc) return launch_cfg<BM, BN, WM, WN, false, true, 1, NSTAGE, MINB>(a, b1, b2, st);
  return launch_cfg<BM, BN, WM, WN, false, false, 1, NSTAGE, MINB>(a, b1, b2, st);
}

inline bool even(long long x) { return (x & 1LL) == 0; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int gemm(const pmr_gemm_t& g0, cudaStream_t st) {
  pmr_gemm_t g = g0;
  PMR_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "dgemm: negative dimension");
  PMR_REQUIRE(g.batch1 >= 1 && g.batch2 >= 1, "dgemm: batch counts must be >= 1");
  if (g.M == 0 || g.N == 0) return 0;
  PMR_REQUIRE(g.C != nullptr, "dgemm: null C");
  PMR_REQUIRE(g.K == 0 || (g.A != nullptr && g.B != nullptr), "dgemm: null operand");
  PMR_REQUIRE(g.a_sm == 1 || g.a_sk == 1, "dgemm: A needs a unit stride (a_sm=%lld a_sk=%lld)",
              g.a_sm, g.a_sk);
  PMR_REQUIRE(g.b_sk == 1 || g.b_sn == 1, "dgemm: B needs a unit stride (b_sk=%lld b_sn=%lld)",
              g.b_sk, g.b_sn);

  if (g.alpha == 0.0) { g.K = 0; g.alpha = 1.0; }  // C = beta*C through the accumulator path

  // Put the long dimension on M (the 128-row tile): C^T = B^T A^T with C's strides swapped.
  if (g.N > g.M && g.flags == 0 && g.C2 == nullptr) {  // (flags != 0: LOWER / k-range / aliasing keep their roles)
    std::swap(g.M, g.N);
    const double* A = g.A;
    const long long a_sm = g.a_sm, a_sk = g.a_sk, a_b1 = g.a_b1, a_b2 = g.a_b2;
    g.A = g.B; g.a_sm = g.b_sn; g.a_sk = g.b_sk; g.a_b1 = g.b_b1; g.a_b2 = g.b_b2;
    g.B = A; g.b_sk = a_sk; g.b_sn = a_sm; g.b_b1 = a_b1; g.b_b2 = a_b2;
    std::swap(g.c_sm, g.c_sn);
  }

  // k-contiguous wins ties (a dimension of extent 1 has a free stride)
  const bool akc = (g.a_sk == 1);
  const bool bkc = (g.b_sk == 1);

  GemmArgs a{};
  a.M = g.M; a.N = g.N; a.K = g.K;
  a.alpha = g.alpha; a.beta = g.beta;
  a.A = g.A; a.B = g.B; a.C = g.C;
  a.a_ld = akc ? g.a_sm : g.a_sk;
  a.b_ld = bkc ? g.b_sn : g.b_sk;
  a.c_sm = g.c_sm; a.c_sn = g.c_sn;
  a.a_b1 = g.a_b1; a.a_b2 = g.a_b2; a.b_b1 = g.b_b1; a.b_b2 = g.b_b2;
  a.c_b1 = g.c_b1; a.c_b2 = g.c_b2;
  a.flags = g.flags;
  a.C2 = g.C2; a.c2_sm = g.c2_sm; a.c2_sn = g.c2_sn; a.c2_b1 = g.c2_b1; a.c2_b2 = g.c2_b2;
  a.a_vec2 = aligned16(g.A) && even(a.a_ld) && even(g.a_b1) && even(g.a_b2);
  a.b_vec2 = aligned16(g.B) && even(a.b_ld) && even(g.b_b1) && even(g.b_b2);
  a.c_vec2 = (g.c_sn == 1) && aligned16(g.C) && even(g.c_sm) && even(g.c_b1) && even(g.c_b2);

  // Tile choice.  128x64 with 3 stages keeps TWO CTAs (2 x 8 warps) resident per SM, so one CTA's
  // store epilogue / pipeline fill overlaps the other's DMMA stream; 128x128 (16 warps, 4 stages)
  // has one CTA per SM.  PMR_GEMM_TILE selects for experiments.
  static const int tile_pref = [] { const char* e = getenv("PMR_GEMM_TILE"); return e ? atoi(e) : 64; }();
  int bn = (g.N > 64) ? tile_pref : (g.N > 32 ? 64 : 32);
  if (g.flags & PMR_GEMM_C_ALIASES_A) {
    // in-place A <- A * B: every CTA must own complete rows, i.e. a single column tile
    PMR_REQUIRE(g.N <= 128, "dgemm: C may alias A only for N <= 128");
    if (g.N > bn) bn = 128;
  }
  // TMA-bulk producer/consumer variant for mn-contiguous operands (see dgemm_bulk_kernel)
  static const bool allow_bulk = [] { const char* e = getenv("PMR_GEMM_NO_BULK"); return !(e && atoi(e)); }();
  // mn-contiguous operands are copied in whole rows, so their extent must be even (16-byte rows)
  const bool a_rows_ok = akc || (g.M % 2) == 0;
  const bool b_rows_ok = bkc || (g.N % 2) == 0;
  // Measured on B200: 1 KiB row copies reach 35.1 TFLOP/s (99 % of cuBLAS DGEMM), but the 128-byte
  // copies a k-contiguous operand needs (one per tile row, 192 per stage) throttle the copy engine
  // (12-17 TFLOP/s), so those layouts stay on the cp.async kernel unless PMR_GEMM_BULK_KC=1.
  static const bool bulk_kc = [] { const char* e = getenv("PMR_GEMM_BULK_KC"); return e && atoi(e); }();
  if (allow_bulk && a.a_vec2 && a.b_vec2 && g.K % BK == 0 && g.K > 0 && a_rows_ok && b_rows_ok &&
      ((!akc && !bkc) || bulk_kc) && !(g.flags & PMR_GEMM_C_ALIASES_A) && g.C2 == nullptr) {
    if (bn == 128) return launch_bulk<128, 128, 32, 32, 4, 1>(a, akc, bkc, g.batch1, g.batch2, st);
    if (bn == 64) return launch_bulk<128, 64, 32, 32, 3, 2>(a, akc, bkc, g.batch1, g.batch2, st);
    return launch_bulk<128, 32, 32, 32, 4, 2>(a, akc, bkc, g.batch1, g.batch2, st);
  }
  // k-contiguous operands on both sides (C -= P P^T trailing updates, S = L_M^T L ...): tensor-map TMA
  // feed, PMR_GEMM_NO_TMA=1 keeps them on the cp.async ring (comparison runs)
  static const bool allow_tma = [] { const char* e = getenv("PMR_GEMM_NO_TMA"); return !(e && atoi(e)); }();
  if (allow_tma && akc && bkc && a.a_vec2 && a.b_vec2 && g.K > 0 && !(g.flags & PMR_GEMM_C_ALIASES_A) &&
      g.C2 == nullptr) {
    int r;
    if (bn == 128) r = launch_tma_cfg<128, 128, 32, 32, 4, 1>(a, g.batch1, g.batch2, st);
    else if (bn == 64) {
      // ring depth: 4 x 24 KiB per CTA by default; PMR_GEMM_TMA_STAGES = 3 / 6 for A/B runs
      static const int stages = [] { const char* e = getenv("PMR_GEMM_TMA_STAGES"); return e ? atoi(e) : TMA_STAGES_64; }();
      if (stages == 3) r = launch_tma_cfg<128, 64, 32, 32, 3, 2>(a, g.batch1, g.batch2, st);
      else if (stages == 6) r = launch_tma_cfg<128, 64, 32, 32, 6, 1>(a, g.batch1, g.batch2, st);
      else r = launch_tma_cfg<128, 64, 32, 32, TMA_STAGES_64, 2>(a, g.batch1, g.batch2, st);
    }
    else r = launch_tma_cfg<128, 32, 32, 32, 4, 2>(a, g.batch1, g.batch2, st);
    if (r <= 0) return r;   // > 0: not expressible as a tensor map, fall through
  }
  if (bn == 128 && (g.flags & PMR_GEMM_C_ALIASES_A))  // in-place panel solve: full rows per CTA,
    return launch_layout<64, 128, 32, 32, 3, 2>(a, akc, bkc, g.batch1, g.batch2, st);  // 2 CTAs / SM
  if (bn == 128) return launch_layout<128, 128, 32, 32, 4, 1>(a, akc, bkc, g.batch1, g.batch2, st);
  if (bn == 64) return launch_layout<128, 64, 32, 32, 3, 2>(a, akc, bkc, g.batch1, g.batch2, st);
  return launch_layout<128, 32, 32, 32, 4, 2>(a, akc, bkc, g.batch1, g.batch2, st);
}

}  // namespace pmr

extern "C" void pmr_profile_enable(int on) {
  pmr::g_prof_on = on != 0;
  if (!on) {
    for (auto& r : pmr::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    pmr::g_prof.clear();
  }
}

// out[cls*4 + {0,1,2,3}] = {launches, total ms, total flops, total algorithmic bytes}, cls < 6:
// 0/1 = 128x128 tile (plain / lower-triangular update), 2/3 = 128x64, 4/5 = 128x32
extern "C" int pmr_profile_fetch(double* out_host, int ncls) {
  for (int k = 0; k < ncls * 4; ++k) out_host[k] = 0.0;
  for (auto& r : pmr::g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return pmr::set_error(PMR_ERR_CUDA, "profile sync");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (r.cls < ncls) {
      out_host[r.cls * 4 + 0] += 1.0;
      out_host[r.cls * 4 + 1] += ms;
      out_host[r.cls * 4 + 2] += r.flops;
      out_host[r.cls * 4 + 3] += r.bytes;
    }
  }
  return 0;
}

extern "C" int pmr_dgemm(const pmr_gemm_t* g, void* stream) {
  if (!g) return pmr::set_error(PMR_ERR_ARG, "pmr_dgemm: null descriptor");
  return pmr::gemm(*g, pmr::as_stream(stream));
}
