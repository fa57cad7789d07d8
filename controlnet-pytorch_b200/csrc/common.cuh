// Shared helpers for libcnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Kernels launched through launch_pdl() may start while their predecessor in the stream is still draining: they run
// their prologue (barrier init, TMEM allocation, descriptor fetch, index math), then pdl_wait() blocks until the
// predecessor grid has completed and its memory is visible.  EVERY global memory access of such a kernel comes after
// pdl_wait(), so stream order is preserved transitively; pdl_trigger() (issued at entry) only lets the NEXT kernel
// begin its own prologue early.  Captured into a CUDA graph these become programmatic dependency edges.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // api.cu: CNB_PDL=1 turns the launch attribute on (off: the device-side calls are no-ops)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace cnb
