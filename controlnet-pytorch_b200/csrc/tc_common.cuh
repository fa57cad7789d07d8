// tcgen05 / TMEM / mbarrier / cp.async PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include "common.cuh"

namespace cnb {

// Latched by a bounded mbarrier wait that expired.  One copy per translation unit (the library is built without
// relocatable device code); every unit exports a reader built on tc_read_clear_error() and cnb_tc_error_flag()
// (conv_tc.cu) ORs them.
static __device__ int g_tc_error = 0;

static inline int tc_read_clear_error() {
  int h = 0;
  if (cudaMemcpyFromSymbol(&h, g_tc_error, sizeof(int)) != cudaSuccess) return -1;
  if (h) {
    int z = 0;
    cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int));
  }
  return h;
}

namespace tc {

constexpr long long WAIT_LIMIT_CYCLES = 4000000000ll;   // ~2 s at 1.9 GHz
// Suspend-time hint of the parked try_wait.  Measured on B200 (tests/probes/lat_probe.cu): without a hint a failing try_wait
// returns after ~6 cycles, so a waiting warp spins through ~135 probes per 1000 cycles and steals issue slots from
// the warps that do the work; with a hint >= 10 us the warp is parked and still wakes ~160 cycles after the arrival.
constexpr uint32_t TRY_WAIT_SUSPEND_NS = 20000u;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the warp is parked instead of re-probing (see TRY_WAIT_SUSPEND_NS)
__device__ __forceinline__ uint32_t mbar_try_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(TRY_WAIT_SUSPEND_NS)
      : "memory");
  return ok;
}
// Bounded wait.  Returns false (and latches *abort) if the phase did not complete within the guard.  Spinning
// variant: lowest wake-up latency, right for kernels with few warps per SM (the convolution pipelines).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
    if (mbar_try_wait(bar, parity)) return true;
    if (*abort) return false;
    if (clock64() - t0 > WAIT_LIMIT_CYCLES) {
      *abort = 1;
      atomicExch(&g_tc_error, 1);
      return false;
    }
  }
}
// Parked variant for kernels with many waiting warps per SM sub-partition (attention_tc05): the first probe is
// inline, the guard loop lives out of line and every probe parks the warp.
static __device__ __noinline__ bool mbar_wait_parked_slow(uint64_t* bar, uint32_t parity, volatile int* abort) {
  const long long t0 = clock64();
  while (true) {
#pragma unroll 1
    for (int spin = 0; spin < 64; ++spin)
      if (mbar_try_wait_parked(bar, parity)) return true;
    if (*abort) return false;
    if (clock64() - t0 > WAIT_LIMIT_CYCLES) {
      *abort = 1;
      atomicExch(&g_tc_error, 1);
      return false;
    }
  }
}
__device__ __forceinline__ bool mbar_wait_parked(uint64_t* bar, uint32_t parity, volatile int* abort) {
  if (mbar_try_wait_parked(bar, parity)) return true;
  return mbar_wait_parked_slow(bar, parity, abort);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <bool HALF>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (HALF) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4 = 1024 B (8 rows)
//   [46,48) version = 1 (sm_100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format (bits 7..12),
// K-major A and B (bits 15,16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool half) {
  uint32_t fmt = half ? 0u : 2u;   // F16F32Format: F16 = 0, BF16 = 1, TF32 = 2
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMA / TMEM helpers shared by the TMA-fed kernels ----
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor for a tile whose rows are RB bytes = one swizzle span
// (cute::UMMA::SmemDescriptor): start >> 4 | LBO (unused) | SBO = 8 rows | version 1 | layout 2 / 4 / 6.
template <int RB>
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr) {
  static_assert(RB == 128 || RB == 64 || RB == 32, "row bytes = swizzle span");
  constexpr uint64_t layout = RB == 128 ? 2 : (RB == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * RB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

}  // namespace tc
}  // namespace cnb
