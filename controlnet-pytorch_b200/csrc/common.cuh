// Shared helpers for libcnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cnb200.h"

namespace cnb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CNB_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      cnb::set_error(__VA_ARGS__);                  \
      return CNB_ERR_BAD_ARG;                       \
    }                                               \
  } while (0)

#define CNB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cnb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CNB_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define CNB_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      cnb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CNB_ERR_CUDA;                                                                   \
    }                                                                                        \
    cnb::count_launch();                                                                     \
  } while (0)

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace cnb
