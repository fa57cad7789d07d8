// TMA-fed, persistent, warp-specialised tcgen05 implicit-GEMM convolution (the wide-channel path of cnb_conv2d).
//
//   GEMM view     M = B*OH*OW output pixels (tile 128 = one UMMA M), N = Cout (tile BN in {16..256}, one CTA usually
//                 owns the full Cout), K = ntaps*Cin walked as (tap, channel chunk) with a chunk = one 128-byte
//                 SWIZZLE_128B row: 64 fp16 (kind::f16) or 32 fp32 (kind::tf32) channels.
//   A operand     cp.async.bulk.tensor.4d ... im2col: ONE instruction gathers the 128 pixels x chunk channels of a tap
//                 straight from the channels-last activation tensor [B, H, W, ldi] into the canonical K-major
//                 SWIZZLE_128B tile (zero fill outside the image and past the last pixel); the tap is the
//                 instruction's {offset_w, offset_h}, the conv stride is the tensor map's traversal stride.
//   B operand     cp.async.bulk.tensor.2d tile {chunk, BN} of the packed weights [Cout][ntaps*Cin].
//   roles         warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer (one elected lane, fp32
//                 accumulators in TMEM, double buffered: 2 x BN columns), warps 2..5 = epilogue (TMEM -> registers ->
//                 per-warp smem slab -> coalesced channels-last stores with + bias + time-embedding row + residual
//                 (+SiLU), fp32 or fp16 out).  S-stage smem ring with full/empty mbarriers; tmem_full/tmem_empty
//                 mbarriers hand accumulators between the MMA warp and the epilogue, so the epilogue of tile i overlaps
//                 the main loop of tile i+1.
//   schedule      persistent: grid = min(#tiles, #SMs), tile = blockIdx.x + k*gridDim.x, n-tile fastest.
//
// Every mbarrier wait is bounded (tc_common.cuh): a broken pipeline raises the hang-guard flag and drains.
#include <cuda.h>   // CUtensorMap + enums only; the encoders are resolved at run time (no link against libcuda)
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace cnb {
namespace tma {

using namespace tc;

constexpr int BM = 128;
constexpr int ROW_BYTES = 128;
constexpr int A_STAGE_BYTES = BM * ROW_BYTES;   // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int SMEM_BUDGET = 227 * 1024;

struct alignas(64) TmaArgs {
  CUtensorMap map_a;
  CUtensorMap map_b;
  const float* bias;
  const float* temb;
  const float* residual;
  void* out;
  int M, OHW, OW;
  int OHf, OWf, oy_mul, oy_add, ox_mul, ox_add;
  int Cout, ldo, out_coff, ldr, res_coff, temb_ld, temb_per_sample, act, out_f16;
  int Cin, ntaps, kchunks;
  int stride, lower_w, lower_h;
  int tiles_m, tiles_n;
  uint32_t tap_off[CNB_MAX_TAPS];   // offset_w | offset_h << 16
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* map, uint64_t* bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(off_w), "h"(off_h)
      : "memory");
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

constexpr int pow2_at_least(int v, int p = 32) { return p >= v ? p : pow2_at_least(v, p * 2); }

template <int BN>
struct Cfg {
  static constexpr int SLAB = (BN % 64 == 0) ? 32 : 16;   // epilogue column slab; BN / SLAB is even where it can be
  static constexpr int SLAB_STRIDE = SLAB + 4;            // floats
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * SLAB_STRIDE * 4;
  static constexpr int B_STAGE_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int S_RAW = (SMEM_BUDGET - 1024 - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;
  static constexpr int S = S_RAW > 8 ? 8 : S_RAW;
  static constexpr int EPI_OFFSET = S * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + BAR_BYTES + 1024;   // + alignment slack
  static constexpr int ACC_STRIDE = pow2_at_least(BN);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static_assert(S >= 2, "pipeline needs at least two stages");
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep 1024-byte alignment (SWIZZLE_128B atoms)");
};

// EPI selects the epilogue: 0 = dense fp32 out, 1 = dense fp32 out + residual, 2 = dense fp16 out, 3 = generic.
template <int BN, bool HALF, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tma_kernel(const __grid_constant__ TmaArgs a) {
  using C = Cfg<BN>;
  constexpr int S = C::S;
  constexpr int KC = HALF ? 64 : 32;                 // channels per stage
  constexpr uint32_t IDESC = make_idesc(BM, BN, HALF);
  constexpr int SLAB = C::SLAB;
  constexpr int SSTR = C::SLAB_STRIDE;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);   // [S]
  uint64_t* empty_bar = full_bar + S;                                       // [S]
  uint64_t* tfull_bar = empty_bar + S;                                      // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                                     // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int nkb = a.ntaps * a.kchunks;
  const int ntiles = a.tiles_m * a.tiles_n;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = tile / a.tiles_n, nt = tile - mt * a.tiles_n;
        const int m0 = mt * BM;
        const int b = m0 / a.OHW;
        const int rem = m0 - b * a.OHW;
        const int oy = rem / a.OW;
        const int ox = rem - oy * a.OW;
        const int w0 = ox * a.stride + a.lower_w, h0 = oy * a.stride + a.lower_h;
        bool ok = true;
        for (int tap = 0; tap < a.ntaps && ok; ++tap) {
          const uint32_t off = a.tap_off[tap];
          for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
            const int s = it % S;
            if (it >= S) ok = mbar_wait(&empty_bar[s], (uint32_t)(((it / S) - 1) & 1), abort_flag);
            if (!ok) break;
            const uint32_t sa = smem_base + (uint32_t)s * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[s], (uint32_t)C::STAGE_BYTES);
            tma_load_im2col_4d(sa, &a.map_a, &full_bar[s], kc * KC, w0, h0, b, (uint16_t)(off & 0xffffu),
                               (uint16_t)(off >> 16));
            tma_load_2d(sa + A_STAGE_BYTES, &a.map_b, &full_bar[s], tap * a.Cin + kc * KC, nt * BN);
          }
        }
        if (!ok) break;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      bool ok = true;
      if (tcount >= 2) ok = mbar_wait(&tempty_bar[acc], (uint32_t)(((tcount >> 1) - 1) & 1), abort_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C::ACC_STRIDE);
      for (int kb = 0; kb < nkb && ok; ++kb, ++it) {
        const int s = it % S;
        ok = mbar_wait(&full_bar[s], (uint32_t)((it / S) & 1), abort_flag);
        tc_fence_after();
        if (lane == 0 && ok) {
          const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE_BYTES;
          const uint64_t adesc = make_desc_sw128(a_addr);
          const uint64_t bdesc = make_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / 32; ++k)
            umma<HALF>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == nkb - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
      if (!ok) break;
    }
    tc_fence_before();
  } else {
    // ============================ epilogue (warps 2..9) ============================
    // Two warps share each TMEM lane quarter (q = warp % 4) and split the accumulator's column slabs by parity.
    // With one warp per scheduler the epilogue is instruction-latency bound, so the dense variants (EPI 0..2) keep
    // the per-row work to LDS.128 (+LDG.128 residual) + 4 FADD + STG and fully unroll it.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int par = ew >> 2;
    float* slab = reinterpret_cast<float*>(smem + C::EPI_OFFSET) + (size_t)ew * 32 * SSTR;
    constexpr int NSLAB = BN / SLAB;
    constexpr int LPR = SLAB / 4;                             // lanes per row in the coalesced pass
    constexpr int RPI = 32 / LPR;                             // rows per iteration
    const int sub_r = lane / LPR, sub_c = (lane % LPR) * 4;
    const int M = a.M, ldo = a.ldo, ldr = a.ldr, tiles_n = a.tiles_n;
    const float* bias = a.bias;
    const float* temb_row = (a.temb && !a.temb_per_sample) ? a.temb : nullptr;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
      const int m_w = mt * BM + q * 32;                       // first GEMM row of this warp
      const int n0 = nt * BN;
      long long pix = -1;
      int b = 0;
      if (EPI == 3) {
        const int m = m_w + lane;
        if (m < M) {
          b = m / a.OHW;
          const int rem = m - b * a.OHW;
          const int oy = rem / a.OW;
          const int ox = rem - oy * a.OW;
          pix = ((long long)b * a.OHf + (oy * a.oy_mul + a.oy_add)) * a.OWf + (ox * a.ox_mul + a.ox_add);
        }
      }
      const int acc = tcount & 1;
      if (!mbar_wait(&tfull_bar[acc], (uint32_t)((tcount >> 1) & 1), abort_flag)) break;
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * C::ACC_STRIDE) + ((uint32_t)(q * 32) << 16);
      if (par >= NSLAB) {                                     // single-slab tiles: the odd warps have nothing to read
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
#pragma unroll 1
      for (int si = par; si < NSLAB; si += 2) {
        const int c0 = si * SLAB;
        uint32_t r[SLAB];
#pragma unroll
        for (int j = 0; j < SLAB / 16; ++j) tmem_ld16_nowait(t_addr + (uint32_t)(c0 + 16 * j), r + 16 * j);
        tmem_ld_wait();
        if (si + 2 >= NSLAB) {            // this warp's last slab: hand its share of the TMEM buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        float4* d = reinterpret_cast<float4*>(slab + (size_t)lane * SSTR);
#pragma unroll
        for (int j = 0; j < SLAB / 4; ++j)
          d[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int n = n0 + c0 + sub_c;                        // < Cout: BN divides Cout
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
        if (temb_row) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(temb_row + n));
          bv.x += t.x; bv.y += t.y; bv.z += t.z; bv.w += t.w;
        }
        if (EPI != 3) {
          // ---- dense output: GEMM row m is output pixel m
          const int rows_left = M - m_w - sub_r;              // row rr + sub_r is valid iff rr < rows_left
          const float* sp = slab + (size_t)sub_r * SSTR + sub_c;
          const size_t o_off = (size_t)(m_w + sub_r) * ldo + a.out_coff + n;
          float4 rv[32 / RPI];
          if (EPI == 1) {
            const float* rp = a.residual + (size_t)(m_w + sub_r) * ldr + a.res_coff + n;
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i)
              if (i * RPI < rows_left) rv[i] = __ldg(reinterpret_cast<const float4*>(rp + (size_t)i * RPI * ldr));
          }
#pragma unroll
          for (int i = 0; i < 32 / RPI; ++i) {
            if (i * RPI < rows_left) {
              float4 o = *reinterpret_cast<const float4*>(sp + i * RPI * SSTR);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              if (EPI == 1) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
              if (EPI == 2) {
                const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + o_off + (size_t)i * RPI * ldo) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
              } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + o_off + (size_t)i * RPI * ldo) = o;
              }
            }
          }
        } else {
          // ---- generic output mapping (ConvTranspose phases, per-sample time embedding, SiLU, fp16 + residual)
#pragma unroll 2
          for (int rr = 0; rr < 32; rr += RPI) {
            const int row = rr + sub_r;
            const long long rp = __shfl_sync(0xffffffffu, pix, row);
            const int rb = __shfl_sync(0xffffffffu, b, row);
            if (rp >= 0) {
              float4 o = *reinterpret_cast<const float4*>(slab + (size_t)row * SSTR + sub_c);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              if (a.temb && a.temb_per_sample) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(a.temb + (size_t)rb * a.temb_ld + n));
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
              }
              if (a.residual) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(a.residual + rp * ldr + a.res_coff + n));
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
              }
              if (a.act == 1) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
              if (a.out_f16) {
                const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.out) + rp * ldo + a.out_coff + n) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
              } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + rp * ldo + a.out_coff + n) = o;
              }
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn g_encode_im2col = nullptr;
static EncodeTiledFn g_encode_tiled = nullptr;
static int g_num_sms = 0;
static int g_driver_version = 0;

static int resolve_driver() {
  if (g_encode_im2col && g_encode_tiled) return CNB_OK;
  cudaDriverEntryPointQueryResult q1, q2;
  void *f1 = nullptr, *f2 = nullptr;
  CNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f1, cudaEnableDefault, &q1));
  CNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f2, cudaEnableDefault, &q2));
  if (!f1 || !f2 || q1 != cudaDriverEntryPointSuccess || q2 != cudaDriverEntryPointSuccess) {
    set_error("conv_tma: the driver does not export cuTensorMapEncodeIm2col / cuTensorMapEncodeTiled");
    return CNB_ERR_UNSUPPORTED;
  }
  int dev = 0;
  CNB_CUDA(cudaGetDevice(&dev));
  CNB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  CNB_CUDA(cudaDriverGetVersion(&g_driver_version));
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f1);
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f2);
  return CNB_OK;
}

template <int BN, bool HALF, int EPI>
static int launch_epi(const TmaArgs& a, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA(cudaFuncSetAttribute(conv_tma_kernel<BN, HALF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
  }
  const int ntiles = a.tiles_m * a.tiles_n;
  const int grid = ntiles < g_num_sms ? ntiles : g_num_sms;
  conv_tma_kernel<BN, HALF, EPI><<<grid, NUM_THREADS, C::TOTAL, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

template <int BN, bool HALF>
static int launch(const TmaArgs& a, cudaStream_t st) {
  const bool dense = a.oy_mul == 1 && a.ox_mul == 1 && a.oy_add == 0 && a.ox_add == 0 && a.OHf * a.OWf == a.OHW;
  const bool simple = dense && a.act == 0 && !(a.temb && a.temb_per_sample);
  if (simple && !a.out_f16 && !a.residual) return launch_epi<BN, HALF, 0>(a, st);
  if (simple && !a.out_f16 && a.residual) return launch_epi<BN, HALF, 1>(a, st);
  if (simple && a.out_f16 && !a.residual) return launch_epi<BN, HALF, 2>(a, st);
  return launch_epi<BN, HALF, 3>(a, st);
}

static int pick_bn(int cout) {
  static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  for (int bn : cand)
    if (cout % bn == 0) return bn;
  return 0;
}

template <bool HALF>
static int dispatch(int bn, const TmaArgs& a, cudaStream_t st) {
  switch (bn) {
    case 256: return launch<256, HALF>(a, st);
    case 192: return launch<192, HALF>(a, st);
    case 128: return launch<128, HALF>(a, st);
    case 96: return launch<96, HALF>(a, st);
    case 64: return launch<64, HALF>(a, st);
    case 48: return launch<48, HALF>(a, st);
    case 32: return launch<32, HALF>(a, st);
    case 16: return launch<16, HALF>(a, st);
  }
  set_error("conv_tma: no tile for Cout");
  return CNB_ERR_UNSUPPORTED;
}

}  // namespace tma

int conv_tma_error_flag() { return tc_read_clear_error(); }

static int g_tma_enabled = -1;

bool conv2d_tma_supported(const cnb_conv_params* p) {
  if (g_tma_enabled < 0) {
    const char* e = getenv("CNB_CONV_TMA");
    g_tma_enabled = e ? atoi(e) : 1;
  }
  if (!g_tma_enabled || p->mode == CNB_MODE_F32) return false;
  const bool half = p->in_dtype == 1;
  const int elt = half ? 2 : 4, kc = half ? 64 : 32;
  if (half && !p->weight_lp) return false;
  if (p->Cin % kc) return false;
  if ((p->ldi * elt) % 16 || (p->in_coff * elt) % 16) return false;
  if (tma::pick_bn(p->Cout) == 0) return false;
  const int oelt = p->out_dtype == 1 ? 2 : 4;
  if (p->ldo % 4 || p->out_coff % 4 || ((size_t)p->out_coff * oelt) % 8) return false;
  if (p->residual && (p->ldr % 4 || p->res_coff % 4)) return false;
  if (p->temb && (p->temb_ld % 4)) return false;
  if (((uintptr_t)p->in | (uintptr_t)p->weight | (uintptr_t)p->weight_lp | (uintptr_t)p->out | (uintptr_t)p->bias |
       (uintptr_t)p->temb | (uintptr_t)p->residual) & 15)
    return false;
  // tap window must be expressible as im2col offsets from a lower corner
  int dxmin = 127, dymin = 127, dxmax = -128, dymax = -128;
  for (int t = 0; t < p->ntaps; ++t) {
    dxmin = p->dx[t] < dxmin ? p->dx[t] : dxmin;
    dymin = p->dy[t] < dymin ? p->dy[t] : dymin;
    dxmax = p->dx[t] > dxmax ? p->dx[t] : dxmax;
    dymax = p->dy[t] > dymax ? p->dy[t] : dymax;
  }
  if (dxmax - dxmin > 255 || dymax - dymin > 255) return false;
  const int uw = dxmin + (p->OW - 1) * p->stride + 1 - p->W, uh = dymin + (p->OH - 1) * p->stride + 1 - p->H;
  if (uw < -128 || uw > 127 || uh < -128 || uh > 127 || p->stride < 1 || p->stride > 8) return false;
  return true;
}

int conv2d_tma(const cnb_conv_params* p, cudaStream_t st) {
  using namespace tma;
  int rc = resolve_driver();
  if (rc != CNB_OK) return rc;
  const bool half = p->in_dtype == 1;
  const int elt = half ? 2 : 4, kc = half ? 64 : 32;
  TmaArgs a;
  memset(&a, 0, sizeof(a));
  int dxmin = 127, dymin = 127;
  for (int t = 0; t < p->ntaps; ++t) {
    dxmin = p->dx[t] < dxmin ? p->dx[t] : dxmin;
    dymin = p->dy[t] < dymin ? p->dy[t] : dymin;
  }
  for (int t = 0; t < p->ntaps; ++t)
    a.tap_off[t] = (uint32_t)(p->dx[t] - dxmin) | ((uint32_t)(p->dy[t] - dymin) << 16);
  const int lower[2] = {dxmin, dymin};
  const int upper[2] = {dxmin + (p->OW - 1) * p->stride + 1 - p->W, dymin + (p->OH - 1) * p->stride + 1 - p->H};

  // ---- activations: [B, H, W, ldi] viewed as (C, W, H, N), im2col box = 128 pixels x kc channels
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->B};
    const cuuint64_t strides[3] = {(cuuint64_t)p->ldi * elt, (cuuint64_t)p->W * p->ldi * elt,
                                   (cuuint64_t)p->H * p->W * p->ldi * elt};
    const cuuint32_t estr[4] = {1, (cuuint32_t)p->stride, (cuuint32_t)p->stride, 1};
    void* base = const_cast<char*>(reinterpret_cast<const char*>(p->in) + (size_t)p->in_coff * elt);
    CUresult r = g_encode_im2col(&a.map_a, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                 base, dims, strides, lower, upper, (cuuint32_t)kc, (cuuint32_t)BM, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeIm2col failed (%d): Cin=%d W=%d H=%d B=%d ldi=%d stride=%d lower=(%d,%d) upper=(%d,%d)",
                (int)r, p->Cin, p->W, p->H, p->B, p->ldi, p->stride, lower[0], lower[1], upper[0], upper[1]);
      return CNB_ERR_CUDA;
    }
    // same small-tensor descriptor fix-up CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp) on drivers <= 13.1
    const size_t bytes = (size_t)p->B * p->H * p->W * p->ldi * elt;
    if (g_driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&a.map_a)[1] &= ~(1llu << 21);
  }
  // ---- packed weights [Cout][K] -> tile {kc, BN}
  const int bn = pick_bn(p->Cout);
  {
    const int K = p->ntaps * p->Cin;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p->Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)K * elt};
    const cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    void* base = const_cast<void*>(half ? p->weight_lp : reinterpret_cast<const void*>(p->weight));
    CUresult r = g_encode_tiled(&a.map_b, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(weights) failed (%d): K=%d Cout=%d", (int)r, K, p->Cout);
      return CNB_ERR_CUDA;
    }
  }
  a.bias = p->bias; a.temb = p->temb; a.residual = p->residual; a.out = p->out;
  a.M = p->B * p->OH * p->OW; a.OHW = p->OH * p->OW; a.OW = p->OW;
  a.OHf = p->OHf; a.OWf = p->OWf;
  a.oy_mul = p->oy_mul; a.oy_add = p->oy_add; a.ox_mul = p->ox_mul; a.ox_add = p->ox_add;
  a.Cout = p->Cout; a.ldo = p->ldo; a.out_coff = p->out_coff; a.ldr = p->ldr; a.res_coff = p->res_coff;
  a.temb_ld = p->temb_ld; a.temb_per_sample = p->temb_per_sample; a.act = p->act; a.out_f16 = p->out_dtype == 1;
  a.Cin = p->Cin; a.ntaps = p->ntaps; a.kchunks = p->Cin / kc;
  a.stride = p->stride; a.lower_w = dxmin; a.lower_h = dymin;
  a.tiles_m = ceil_div(a.M, BM); a.tiles_n = p->Cout / bn;
  return half ? dispatch<true>(bn, a, st) : dispatch<false>(bn, a, st);
}

}  // namespace cnb
