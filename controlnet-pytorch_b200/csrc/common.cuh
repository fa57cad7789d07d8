// Shared helpers for libcnb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// fp16 activation stream, overflow policy: SATURATE.  An activation beyond the fp16 range (65504) is stored as +-65504,
// never as inf - an inf in the residual stream turns the next GroupNorm's statistics into NaN for the whole sample, a
// saturated value only distorts its own group (tests/test_models_gpu.py::test_f16_mode_activation_range).  Two FMNMX per
// pair on a store-bound epilogue: below measurement noise on the step.
__device__ __forceinline__ __half2 h2_sat(float x, float y) {
  // clamp AFTER the conversion (two packed min / max instead of four scalar ones): round-to-nearest maps everything
  // below 65520 to <= 65504 already, and what it maps to +-inf is exactly what the clamp brings back to +-65504
  const __half2 lim = __floats2half2_rn(65504.f, 65504.f);
  return __hmax2(__hmin2(__floats2half2_rn(x, y), lim), __hneg2(lim));
}


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Function attributes (MaxDynamicSharedMemorySize) are PER DEVICE: a process that touches cuda:1 after cuda:0 must set
// them again there.  `static DeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)` at each launch site.
// (Device PROPERTIES cached process-wide - SM count, tcgen05 availability - assume the homogeneous GPUs of one box.)
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Kernels launched through launch_pdl() may start while their predecessor in the stream is still draining: they run
// their prologue (barrier init, TMEM allocation, descriptor fetch, index math), then pdl_wait() blocks until the
// predecessor grid has completed and its memory is visible.  EVERY global memory access of such a kernel comes after
// pdl_wait(), so stream order is preserved transitively; pdl_trigger() (issued at entry) only lets the NEXT kernel
// begin its own prologue early.  Captured into a CUDA graph these become programmatic dependency edges.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// api.cu.  `work` = output elements of the launch.  CNB_PDL=0 never, =1 always; default (auto): only for launches that
// write at most CNB_PDL_WORK (4 Mi) elements.  Measured on the replayed MNIST step graph: always-on gains 7.6 % at
// B = 16, 5.3 % at 64, 1.2 % at 256, nothing at 512 and LOSES 1 % at 1024 (12.74 vs 12.61 ms; also 12.82 ms when
// only the 7x7 layers' launches at B = 1024, 6-13 Mi elements each, carry the attribute) - a dependent's CTAs parked
// at griddepcontrol.wait take slots the other capture stream's kernels would have filled - so the attribute is set
// only where the launch is short enough (under ~10 us) for its ramp-up and tail to matter.
bool pdl_enabled(long long work);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(long long work, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(work) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace cnb
