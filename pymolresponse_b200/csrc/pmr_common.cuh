// Shared host/device helpers for libpmr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "pmr_b200.h"

namespace pmr {

// per-thread error text + launch counter (the only state the library keeps)
char* error_buffer();
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define PMR_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return pmr::set_error(PMR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define PMR_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return pmr::set_error(PMR_ERR_CUDA, "launch of %s failed: %s", name,                \
                            cudaGetErrorString(_e));                                      \
    pmr::count_launch();                                                                  \
  } while (0)

#define PMR_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) return pmr::set_error(PMR_ERR_ARG, __VA_ARGS__);                         \
  } while (0)

#define PMR_TRY(expr)                                                                     \
  do {                                                                                    \
    int _s = (expr);                                                                      \
    if (_s != 0) return _s;                                                               \
  } while (0)

// internal C++ entry used by the other translation units
int gemm(const pmr_gemm_t& g, cudaStream_t stream);

// convenience: plain (unbatched) GEMM descriptor
inline pmr_gemm_t make_gemm(int M, int N, int K, double alpha, const double* A, long long a_sm,
                            long long a_sk, const double* B, long long b_sk, long long b_sn,
                            double beta, double* C, long long c_sm, long long c_sn) {
  pmr_gemm_t g{};
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
  g.A = A; g.a_sm = a_sm; g.a_sk = a_sk;
  g.B = B; g.b_sk = b_sk; g.b_sn = b_sn;
  g.C = C; g.c_sm = c_sm; g.c_sn = c_sn;
  g.batch1 = 1; g.batch2 = 1;
  g.flags = 0;
  return g;
}

}  // namespace pmr
