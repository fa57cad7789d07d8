// Latency probe: tcgen05.mma -> commit -> mbarrier visible; mbarrier arrive -> try_wait wake; tcgen05.ld.
#include <cstdio>
#include <cuda.h>
#include <cuda_fp16.h>
#include "../../controlnet-pytorch_b200/csrc/tc_common.cuh"
using namespace cnb;
using namespace cnb::tc;

__device__ __forceinline__ uint32_t try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint)
      : "memory");
  return ok;
}
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma, int N, uint32_t hint) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile long long* stamp = reinterpret_cast<volatile long long*>(bars + 16);
  volatile int* abortf = reinterpret_cast<volatile int*>(slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    *abortf = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async();
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // ---- A: mma issue -> commit -> wait
  if (warp == 0 && lane == 0) {
    const uint64_t adesc = make_desc_kmajor<32>(sbase), bdesc = make_desc_kmajor<32>(sbase + 8192);
    const uint32_t idesc = make_idesc(128, N, true);
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
      for (int k = 0; k < n_mma; ++k) umma<true>(tmem, adesc, bdesc, idesc, k ? 1u : 0u);
      umma_commit(&bars[0]);
      const long long t1 = clock64();
      while (!mbar_try_wait(&bars[0], rep & 1)) {}
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  __syncthreads();
  // ---- B: arrive (warp 1) -> wake (warp 2 waiting in try_wait); 8 reps with a delay so the waiter is parked
  if (warp == 1 && lane == 0) {
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
      while (clock64() - t0 < 20000) {}
      stamp[rep] = clock64();
      __threadfence_block();
      mbar_arrive(&bars[1]);
      // wait for the echo
      while (!mbar_try_wait(&bars[2], rep & 1)) {}
    }
  } else if (warp == 2 && lane == 0) {
    for (int rep = 0; rep < 8; ++rep) {
      int tries = 0;
      if (hint) { while (!try_wait_hint(&bars[1], rep & 1, hint)) ++tries; }
      else { while (!mbar_try_wait(&bars[1], rep & 1)) ++tries; }
      const long long t = clock64();
      out[16 + rep * 2] = t - stamp[rep];
      out[16 + rep * 2 + 1] = tries;
      mbar_arrive(&bars[2]);
    }
  }
  __syncthreads();
  // ---- C: tcgen05.ld x32 latency (warp 0)
  if (warp == 0) {
    uint32_t r[32];
    for (int rep = 0; rep < 4; ++rep) {
      const long long t0 = clock64();
      tmem_ld16_nowait(tmem, r);
      tmem_ld16_nowait(tmem + 16, r + 16);
      tmem_ld_wait();
      const long long t1 = clock64();
      uint32_t acc = 0;
      for (int i = 0; i < 32; ++i) acc ^= r[i];
      if (lane == 0) out[32 + rep] = (t1 - t0) + (acc == 0x12345 ? 1 : 0);
    }
  }
  // ---- D: test_wait-style single probe cost: time 100 failing try_waits
  if (warp == 3 && lane == 0) {
    const long long t0 = clock64();
    int c = 0;
    if (hint) { for (int i = 0; i < 100; ++i) c += try_wait_hint(&bars[5], 0, hint); }
    else { for (int i = 0; i < 100; ++i) c += mbar_try_wait(&bars[5], 0); }
    const long long t1 = clock64();
    out[40] = (t1 - t0) + c * 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8); 
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048);
  for (uint32_t hint : {0u, 100u, 1000u, 10000u, 1000000u}) for (int N : {64}) for (int n_mma : {1, 2}) {
    printf("hint=%u ", hint);
    cudaMemset(d, 0, 64 * 8);
    probe<<<1, 128, 65536 + 2048>>>(d, n_mma, N, hint);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d n_mma=%d err=%d\n  A issue/total:", N, n_mma, (int)e);
    for (int i = 0; i < 8; ++i) printf(" %lld/%lld", h[2 * i], h[2 * i + 1]);
    printf("\n  B wake(cycles)/tries:");
    for (int i = 0; i < 8; ++i) printf(" %lld/%lld", h[16 + 2 * i], h[17 + 2 * i]);
    printf("\n  C ldtm x32:");
    for (int i = 0; i < 4; ++i) printf(" %lld", h[32 + i]);
    printf("\n  D 100 failing try_wait: %lld cycles\n", h[40]);
  }
  return 0;
}
