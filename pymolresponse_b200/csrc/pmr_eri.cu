// AO two-electron integrals with 8-fold permutational symmetry: only the unique part crosses PCIe.
//
// (mu nu|la si) = (nu mu|la si) = (mu nu|si la) = (la si|mu nu) for real orbitals, which is what every
// caller of the reference hands to AO2MO (pyscf int2e / psi4 ao_eri, reference solvers.py:86-127).
// With P = mu*n + nu the tensor is a symmetric n^2 x n^2 matrix whose rows (mu, nu) and (nu, mu)
// are equal.  pmr_eri_upload_s8 copies, for every mu, the rows (mu, 0..mu) truncated after column
// (mu, mu) -- one strided DMA per mu, about a third of the n^4 doubles -- and pmr_eri_complete_s8
// rebuilds the full tensor in HBM with two streaming kernels:
//   A. row mirror:       row (nu, mu) <- row (mu, nu), nu < mu, over the uploaded prefix;
//   B. transpose mirror: entry (Q, P) <- entry (P, Q), Q < P (64x64 tiles through shared memory).
// After A every row P holds at least its columns Q <= P, so B reads only defined entries.
//
// layout 1 ("box", about a sixth of the doubles): for every la one strided DMA of the entries
// (mu >= la, all nu | la, si <= la); completion in three passes, m(P) = max(mu, nu):
//   1. in every uploaded row (mu, nu): (si, la) <- (la, si) for si < la <= mu;
//   2. row (mu, nu), mu < nu  <-  row (nu, mu), columns (la, si) in [0, nu]^2;
//      now every row P holds the columns Q with m(Q) <= m(P);
//   3. entry (P, Q) with m(Q) > m(P)  <-  entry (Q, P)   (64x64 tiles through shared memory).
//
// layout 2 ("tail", the same sixth of the doubles in LONG runs): for every la one strided DMA of the
// rows (la, si <= la), columns from (la, 0) to the end -- the pair-swapped image of the box, i.e. the
// entries (P, Q) with p2 <= p1 <= q1.  The narrow rows of the box cost the DMA a fixed 6.5 ns each
// (8.4 M rows at n = 256: 160 ms for 5.8 GB); here the same bytes move at the full PCIe rate (105 ms).
// Completion, with m(P) = max(p1, p2):
//   1. row (b, a), b < a  <-  row (a, b) on the columns from (a, 0) on; every row P now holds the
//      columns Q with q1 >= m(P);
//   2. in every row: (d, c) <- (c, d) for d < m(P) <= c; now the columns with m(Q) >= m(P);
//   3. entry (P, Q) with m(Q) < m(P)  <-  entry (Q, P).
#include <algorithm>

#include "pmr_common.cuh"

namespace {

using namespace pmr;

__global__ void __launch_bounds__(256) eri_row_mirror_kernel(double* __restrict__ I, int n) {
  const int mu = blockIdx.z, nu = blockIdx.y;
  if (nu >= mu) return;
  const long long n2 = (long long)n * n;
  const long long width = (long long)mu * n + mu + 1;
  const double* src = I + ((long long)mu * n + nu) * n2;
  double* dst = I + ((long long)nu * n + mu) * n2;
  const long long base = (long long)blockIdx.x * (256 * 8);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const long long c = base + u * 256 + threadIdx.x;
    if (c < width) dst[c] = src[c];
  }
}

constexpr int TT = 64;

__global__ void __launch_bounds__(256) eri_transpose_mirror_kernel(double* __restrict__ I,
                                                                   long long P) {
  // destination tile (tr, tc) with tc >= tr receives the transpose of source tile (tc, tr)
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc < tr) return;
  __shared__ double tile[TT][TT + 1];
  const long long r0 = (long long)tc * TT, c0 = (long long)tr * TT;   // source tile origin
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;             // 64 x 4
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = r0 + r, cc = c0 + tx;
    if (rr < P && cc < P) tile[r][tx] = I[rr * P + cc];
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = c0 + r, cc = r0 + tx;     // destination (row in tr band, col in tc band)
    if (rr < P && cc < P && cc > rr) I[rr * P + cc] = tile[tx][r];
  }
}

// ---- layout 1 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eri_box_pass1_kernel(double* __restrict__ I, int n) {
  // one block per row (mu, nu): symmetrise the leading (mu+1)^2 block of its n x n matrix
  const int mu = blockIdx.y, nu = blockIdx.x;
  double* row = I + ((long long)mu * n + nu) * (long long)n * n;
  const int m = mu + 1;
  for (int idx = threadIdx.x; idx < m * m; idx += 256) {
    const int si = idx / m, la = idx - si * m;       // destination (si, la): write coalesced in la
    if (la > si) row[(long long)si * n + la] = row[(long long)la * n + si];
  }
}

__global__ void __launch_bounds__(256) eri_box_pass2_kernel(double* __restrict__ I, int n) {
  // row (mu, nu) with mu < nu  <-  row (nu, mu) on the block [0, nu]^2
  const int mu = blockIdx.y, nu = blockIdx.x;
  if (mu >= nu) return;
  const long long n2 = (long long)n * n;
  const double* src = I + ((long long)nu * n + mu) * n2;
  double* dst = I + ((long long)mu * n + nu) * n2;
  const int m = nu + 1;
  for (int idx = threadIdx.x; idx < m * m; idx += 256) {
    const int la = idx / m, si = idx - la * m;
    dst[(long long)la * n + si] = src[(long long)la * n + si];
  }
}

__device__ __forceinline__ int pair_max(long long P, int n) {
  const int a = (int)(P / n), b = (int)(P - (long long)a * n);
  return a > b ? a : b;
}

__global__ void __launch_bounds__(256) eri_box_pass3_kernel(double* __restrict__ I, int n,
                                                            long long P) {
  // destination tile (rows r0.., cols c0..): entries with m(col) > m(row) take the transposed one
  __shared__ double tile[TT][TT + 1];
  const long long r0 = (long long)blockIdx.y * TT, c0 = (long long)blockIdx.x * TT;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  // any work in this tile?  (the largest m over its columns against the smallest over its rows)
  const long long cc = c0 + tx;
  const int mcol = cc < P ? pair_max(cc, n) : -1;
  int any = 0;
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = r0 + r;
    if (rr < P && mcol > pair_max(rr, n)) any = 1;
  }
  if (!__syncthreads_or(any)) return;
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long sr = c0 + r, sc = r0 + tx;       // source tile = transposed position
    if (sr < P && sc < P) tile[r][tx] = I[sr * P + sc];
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = r0 + r;
    if (rr < P && cc < P && mcol > pair_max(rr, n)) I[rr * P + cc] = tile[tx][r];
  }
}

// ---- layout 2 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eri_tail_pass1_kernel(double* __restrict__ I, int n) {
  const int a = blockIdx.z, b = blockIdx.y;
  if (b >= a) return;
  const long long n2 = (long long)n * n;
  const long long start = (long long)a * n;
  const double* src = I + ((long long)a * n + b) * n2;
  double* dst = I + ((long long)b * n + a) * n2;
  const long long base = start + (long long)blockIdx.x * (256 * 8);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const long long c = base + u * 256 + threadIdx.x;
    if (c < n2) dst[c] = src[c];
  }
}

__global__ void __launch_bounds__(256) eri_tail_pass2_kernel(double* __restrict__ I, int n) {
  // one block per row P = (p1, p2): (d, c) <- (c, d) for d < m <= c
  const int p1 = blockIdx.y, p2 = blockIdx.x;
  const int m = p1 > p2 ? p1 : p2;
  const int w = n - m;
  if (m == 0 || w == 0) return;
  double* row = I + ((long long)p1 * n + p2) * (long long)n * n;
  for (int idx = threadIdx.x; idx < m * w; idx += 256) {
    const int d = idx / w, c = m + (idx - d * w);       // destination (d, c): coalesced in c
    row[(long long)d * n + c] = row[(long long)c * n + d];
  }
}

__global__ void __launch_bounds__(256) eri_tail_pass3_kernel(double* __restrict__ I, int n,
                                                             long long P) {
  // destination tile (rows r0.., cols c0..): entries with m(col) < m(row) take the transposed one
  __shared__ double tile[TT][TT + 1];
  const long long r0 = (long long)blockIdx.y * TT, c0 = (long long)blockIdx.x * TT;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const long long cc = c0 + tx;
  const int mcol = cc < P ? pair_max(cc, n) : (1 << 30);
  int any = 0;
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = r0 + r;
    if (rr < P && mcol < pair_max(rr, n)) any = 1;
  }
  if (!__syncthreads_or(any)) return;
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long sr = c0 + r, sc = r0 + tx;
    if (sr < P && sc < P) tile[r][tx] = I[sr * P + sc];
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < TT; r += 4) {
    const long long rr = r0 + r;
    if (rr < P && cc < P && mcol < pair_max(rr, n)) I[rr * P + cc] = tile[tx][r];
  }
}

}  // namespace

extern "C" long long pmr_eri_s8_bytes(int n, int layout) {
  long long total = 0;
  if (layout == 1 || layout == 2) {
    for (long long la = 0; la < n; ++la) total += (n - la) * n * (la + 1) * 8;
  } else {
    for (long long mu = 0; mu < n; ++mu) total += (mu + 1) * (mu * n + mu + 1) * 8;
  }
  return total;
}

// Narrow strided rows cost a fixed time per row on top of the bytes (measured on B200, n = 256:
// 5.79 GB in 160 ms = 36 GB/s against 55 GB/s for long runs).  The DMAs can be spread over side
// streams measured 1..8 streams: no gain, the limit is the PCIe read
// of short host rows, not the copy engine -- so the default is the caller's stream alone.
constexpr int kMaxUploadStreams = 8;
static const int g_upload_streams = 1;  // (1..8 side streams measured: no gain, the PCIe read is the limit)

struct UploadStreams {
  cudaStream_t s[kMaxUploadStreams] = {};
  cudaEvent_t fork = nullptr, join[kMaxUploadStreams] = {};
  int device = -1;
};
static thread_local UploadStreams g_up;

static int upload_streams_ready() {
  int dev = 0;
  PMR_CHECK_CUDA(cudaGetDevice(&dev));
  if (g_up.device == dev) return 0;
  for (int i = 0; i < kMaxUploadStreams; ++i) {
    PMR_CHECK_CUDA(cudaStreamCreateWithFlags(&g_up.s[i], cudaStreamNonBlocking));
    PMR_CHECK_CUDA(cudaEventCreateWithFlags(&g_up.join[i], cudaEventDisableTiming));
  }
  PMR_CHECK_CUDA(cudaEventCreateWithFlags(&g_up.fork, cudaEventDisableTiming));
  g_up.device = dev;
  return 0;
}


extern "C" int pmr_eri_upload_s8(const double* host, double* dev, int n, int layout, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 0 && layout >= 0 && layout <= 2, "eri_upload_s8: bad n or layout");
  if (n == 0) return 0;
  PMR_REQUIRE(host && dev, "eri_upload_s8: null pointer");
  const size_t n2 = (size_t)n * n;
  if (layout == 2) {
    for (size_t la = 0; la < (size_t)n; ++la) {
      const size_t off = la * n * n2 + la * n;        // row (la, 0), column (la, 0)
      PMR_CHECK_CUDA(cudaMemcpy2DAsync(dev + off, n2 * sizeof(double), host + off,
                                       n2 * sizeof(double), (n - la) * n * sizeof(double), la + 1,
                                       cudaMemcpyHostToDevice, st));
    }
    return 0;
  }
  if (layout == 1) {
    const int ns = g_upload_streams;
    if (ns > 1) {
      PMR_TRY(upload_streams_ready());
      PMR_CHECK_CUDA(cudaEventRecord(g_up.fork, st));
      for (int i = 0; i < ns; ++i) PMR_CHECK_CUDA(cudaStreamWaitEvent(g_up.s[i], g_up.fork, 0));
    }
    for (size_t la = 0; la < (size_t)n; ++la) {
      const size_t off = la * n * n2 + la * n;        // row (la, 0), column (la, 0)
      cudaStream_t cs = ns > 1 ? g_up.s[la % ns] : st;
      PMR_CHECK_CUDA(cudaMemcpy2DAsync(dev + off, n2 * sizeof(double), host + off,
                                       n2 * sizeof(double), (la + 1) * sizeof(double),
                                       (n - la) * n, cudaMemcpyHostToDevice, cs));
    }
    if (ns > 1)
      for (int i = 0; i < ns; ++i) {
        PMR_CHECK_CUDA(cudaEventRecord(g_up.join[i], g_up.s[i]));
        PMR_CHECK_CUDA(cudaStreamWaitEvent(st, g_up.join[i], 0));
      }
    return 0;
  }
  for (size_t mu = 0; mu < (size_t)n; ++mu) {
    const size_t width = (mu * n + mu + 1) * sizeof(double);
    const size_t off = mu * n * n2;
    PMR_CHECK_CUDA(cudaMemcpy2DAsync(dev + off, n2 * sizeof(double), host + off, n2 * sizeof(double),
                                     width, mu + 1, cudaMemcpyHostToDevice, st));
  }
  return 0;
}

extern "C" int pmr_eri_complete_s8(double* dev, int n, int layout, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 0 && n <= 2047 && layout >= 0 && layout <= 2,
              "eri_complete_s8: n out of range or bad layout");
  if (n <= 0) return 0;
  PMR_REQUIRE(dev, "eri_complete_s8: null pointer");
  const long long P = (long long)n * n;
  const unsigned nt = (unsigned)((P + TT - 1) / TT);
  if (layout == 2) {
    if (n > 1) {
      const long long wmax = (long long)(n - 1) * n;   // widest tail: a = 1
      dim3 grid((unsigned)((wmax + 2047) / 2048), (unsigned)n, (unsigned)n);
      eri_tail_pass1_kernel<<<grid, 256, 0, st>>>(dev, n);
      PMR_CHECK_LAUNCH("eri_tail_pass1_kernel");
      eri_tail_pass2_kernel<<<dim3(n, n), 256, 0, st>>>(dev, n);
      PMR_CHECK_LAUNCH("eri_tail_pass2_kernel");
    }
    eri_tail_pass3_kernel<<<dim3(nt, nt), 256, 0, st>>>(dev, n, P);
    PMR_CHECK_LAUNCH("eri_tail_pass3_kernel");
    return 0;
  }
  if (layout == 1) {
    if (n > 1) {
      eri_box_pass1_kernel<<<dim3(n, n), 256, 0, st>>>(dev, n);
      PMR_CHECK_LAUNCH("eri_box_pass1_kernel");
      eri_box_pass2_kernel<<<dim3(n, n), 256, 0, st>>>(dev, n);
      PMR_CHECK_LAUNCH("eri_box_pass2_kernel");
    }
    eri_box_pass3_kernel<<<dim3(nt, nt), 256, 0, st>>>(dev, n, P);
    PMR_CHECK_LAUNCH("eri_box_pass3_kernel");
    return 0;
  }
  if (n > 1) {
    const long long wmax = (long long)(n - 1) * n + n;
    dim3 grid((unsigned)((wmax + 2047) / 2048), (unsigned)n, (unsigned)n);
    eri_row_mirror_kernel<<<grid, 256, 0, st>>>(dev, n);
    PMR_CHECK_LAUNCH("eri_row_mirror_kernel");
  }
  eri_transpose_mirror_kernel<<<dim3(nt, nt), 256, 0, st>>>(dev, P);
  PMR_CHECK_LAUNCH("eri_transpose_mirror_kernel");
  return 0;
}

// ---- slabs (rows = any set of (mu, nu) pairs, e.g. the nu-shard of one rank) ----------------------
// Only (la si| = (si la| can be used inside a slab: for every la one strided DMA of the columns
// (la, si <= la) of all rows, then each row's n x n matrix is symmetrised in place.
namespace {
__global__ void __launch_bounds__(256) eri_ket_mirror_kernel(double* __restrict__ I, int n) {
  double* row = I + (long long)blockIdx.x * n * n;
  for (int idx = threadIdx.x; idx < n * n; idx += 256) {
    const int si = idx / n, la = idx - si * n;
    if (la > si) row[(long long)si * n + la] = row[(long long)la * n + si];
  }
}
}  // namespace

extern "C" long long pmr_eri_ket_s2_bytes(long long rows, int n) {
  return rows * ((long long)n * (n + 1) / 2) * 8;
}

extern "C" int pmr_eri_upload_ket_s2(const double* host, double* dev, long long rows, int n,
                                     void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(rows >= 0 && n >= 0, "eri_upload_ket_s2: bad sizes");
  if (rows == 0 || n == 0) return 0;
  PMR_REQUIRE(host && dev, "eri_upload_ket_s2: null pointer");
  const size_t n2 = (size_t)n * n;
  for (size_t la = 0; la < (size_t)n; ++la) {
    const size_t off = la * n;
    PMR_CHECK_CUDA(cudaMemcpy2DAsync(dev + off, n2 * sizeof(double), host + off, n2 * sizeof(double),
                                     (la + 1) * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice,
                                     st));
  }
  return 0;
}

extern "C" int pmr_eri_complete_ket_s2(double* dev, long long rows, int n, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(rows >= 0 && rows <= 2147483647LL && n >= 0, "eri_complete_ket_s2: bad sizes");
  if (rows == 0 || n <= 1) return 0;
  PMR_REQUIRE(dev, "eri_complete_ket_s2: null pointer");
  eri_ket_mirror_kernel<<<(unsigned)rows, 256, 0, st>>>(dev, n);
  PMR_CHECK_LAUNCH("eri_ket_mirror_kernel");
  return 0;
}
