// Cost of the synchronisation primitives used per tile by the attention softmax warps (single warp, idle SM).
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include "../../controlnet-pytorch_b200/csrc/tc_common.cuh"
using namespace cnb;
using namespace cnb::tc;
__device__ __forceinline__ uint32_t try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
#define TIME(idx, N, stmt) { __syncwarp(); const long long t0 = clock64(); _Pragma("unroll 1") for (int i = 0; i < N; ++i) { stmt; } const long long t1 = clock64(); if (lane == 0) out[idx] = (t1 - t0) / N; }
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  __shared__ __align__(16) uint64_t bars[8];
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint8_t buf[4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&slot, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    uint32_t acc = 0;
    if (lane == 0) mbar_arrive(&bars[0]);    // phase 0 of bar 0 complete
    __syncwarp();
    TIME(0, 16, acc += mbar_try_wait(&bars[0], 0));       // successful try_wait with hint
    TIME(1, 16, acc += try_wait_nohint(&bars[0], 0));      // successful try_wait without hint
    TIME(2, 16, acc += test_wait(&bars[0], 0));            // successful test_wait
    TIME(3, 16, if (lane == 0) mbar_arrive(&bars[1]));     // arrive (count 1 -> completes every time)
    TIME(4, 16, fence_proxy_async());
    TIME(5, 16, tc_fence_after());
    TIME(6, 16, tc_fence_before());
    TIME(7, 16, __syncwarp());
    uint32_t r[32];
    TIME(8, 16, tmem_ld16_nowait(tmem, r); tmem_ld16_nowait(tmem + 16, r + 16); tmem_ld_wait(); acc += r[0] + r[31]);
    TIME(9, 16, *reinterpret_cast<uint4*>(buf + ((lane * 128 + (i & 7) * 16) & 4095)) = make_uint4(acc, i, 2, 3));
    TIME(10, 16, *reinterpret_cast<uint4*>(buf + ((lane * 128 + (i & 7) * 16) & 4095)) = make_uint4(acc, i, 2, 3); fence_proxy_async());
    TIME(11, 16, acc += __any_sync(0xffffffffu, acc > i));
    TIME(12, 16, acc += (uint32_t)clock64());
    if (acc == 0x12345678) out[63] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}
int main() {
  long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
  probe<<<1, 128>>>(d);
  printf("err=%d\n", (int)cudaDeviceSynchronize());
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* nm[13] = {"try_wait(hint) success", "try_wait success", "test_wait success", "arrive (lane 0)", "fence.proxy.async", "tcgen05.fence after",
                        "tcgen05.fence before", "syncwarp", "LDTM x32 + wait", "STS.128", "STS.128 + fence.proxy.async", "any_sync", "clock64"};
  for (int i = 0; i < 13; ++i) printf("%-30s %lld cycles\n", nm[i], h[i]);
  return 0;
}
