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

}  // namespace

extern "C" long long pmr_eri_s8_bytes(int n) {
  long long total = 0;
  for (long long mu = 0; mu < n; ++mu) total += (mu + 1) * (mu * n + mu + 1) * 8;
  return total;
}

extern "C" int pmr_eri_upload_s8(const double* host, double* dev, int n, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 0, "eri_upload_s8: bad n");
  if (n == 0) return 0;
  PMR_REQUIRE(host && dev, "eri_upload_s8: null pointer");
  const size_t n2 = (size_t)n * n;
  for (size_t mu = 0; mu < (size_t)n; ++mu) {
    const size_t width = (mu * n + mu + 1) * sizeof(double);
    const size_t off = mu * n * n2;
    PMR_CHECK_CUDA(cudaMemcpy2DAsync(dev + off, n2 * sizeof(double), host + off, n2 * sizeof(double),
                                     width, mu + 1, cudaMemcpyHostToDevice, st));
  }
  return 0;
}

extern "C" int pmr_eri_complete_s8(double* dev, int n, void* stream) {
  cudaStream_t st = as_stream(stream);
  PMR_REQUIRE(n >= 0 && n <= 2047, "eri_complete_s8: n out of range");
  if (n <= 0) return 0;
  PMR_REQUIRE(dev, "eri_complete_s8: null pointer");
  const long long P = (long long)n * n;
  if (n > 1) {
    const long long wmax = (long long)(n - 1) * n + n;
    dim3 grid((unsigned)((wmax + 2047) / 2048), (unsigned)n, (unsigned)n);
    eri_row_mirror_kernel<<<grid, 256, 0, st>>>(dev, n);
    PMR_CHECK_LAUNCH("eri_row_mirror_kernel");
  }
  const unsigned nt = (unsigned)((P + TT - 1) / TT);
  eri_transpose_mirror_kernel<<<dim3(nt, nt), 256, 0, st>>>(dev, P);
  PMR_CHECK_LAUNCH("eri_transpose_mirror_kernel");
  return 0;
}
