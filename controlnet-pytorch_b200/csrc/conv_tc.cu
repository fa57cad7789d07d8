// tcgen05 implicit-GEMM convolution (CNB_MODE_F16): the tensor-core path of cnb_conv2d on sm_100a.
//
//   GEMM view        M = B*OH*OW output pixels (tile 128 = one UMMA M),  N = Cout (tile BN in {16,...,256}),
//                    K = ntaps*Cin (tile = one 128-byte swizzle row: 32 fp32 (kind::tf32) or 64 fp16 (kind::f16)).
//   operand types    fp32 activations run as kind::tf32; activations that GroupNorm already emitted as fp16
//                    (in_dtype = 1; 10-bit mantissa like tf32, half the bytes, twice the MMA rate) run as kind::f16
//                    against an fp16 copy of the packed weights.  Accumulation is always fp32 in TMEM.
//   operands         A (im2col gather of the channels-last activations, zero outside the image) and W ([N][K]) are
//                    staged into shared memory in the canonical K-major SWIZZLE_128B UMMA layout
//                    (row r at r*128 B, 16-byte chunk c stored at chunk c ^ (r & 7)) by 128 producer threads with
//                    16-byte cp.async (zero-fill for padding / K tail), S-stage ring, mbarrier full/empty pairs.
//   math             one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) x4 per stage,
//                    fp32 accumulator in TMEM (BN columns), tcgen05.commit releases smem stages / signals the epilogue.
//   epilogue         the 4 producer warps read their 32 TMEM lanes with tcgen05.ld.32x32b.x16 and apply
//                    + bias + time-embedding row + residual (+SiLU) before the channels-last store.
//
// Every mbarrier wait is bounded by a clock64() guard: on expiry the CTA raises g_tc_error and drains instead of
// hanging the GPU (cnb_tc_error_flag() reports it).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace cnb {

int conv_tma_error_flag();        // conv_tma.cu
int attention_tmem_error_flag();  // attention_tmem.cu

namespace tc {

constexpr int BM = 128;
constexpr int BKB = 128;                 // bytes of K per stage row (one SWIZZLE_128B row)
constexpr int A_STAGE_BYTES = BM * BKB;  // 16 KB
constexpr int NUM_PRODUCERS = 128;

struct TcArgs {
  cnb_conv_params p;
  int M, K;
};

template <int BN, int S>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BKB;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = S * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 128 * 12 + 1024;   // + barriers/tmem slot + row tables + 1 KB alignment slack
};

template <int BN, int S, bool HALF>
__global__ void __launch_bounds__(160, 1)
conv_igemm_tc_kernel(const __grid_constant__ TcArgs a) {
  using L = SmemLayout<BN, S>;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = make_idesc(BM, BN, HALF);
  constexpr int ELT = HALF ? 2 : 4;          // bytes per operand element
  constexpr int EPC = 16 / ELT;              // elements per 16-byte chunk
  constexpr int BKE = BKB / ELT;             // K elements per stage
  constexpr int AROWS = BM / 16;             // A rows gathered per producer thread (8)
  constexpr int BROWS = (BN + 15) / 16;      // W rows per producer thread
  constexpr int OSTR = BN + 4;               // epilogue staging row stride (floats)
  static_assert(BM * OSTR * 4 <= S * L::STAGE_BYTES, "epilogue staging must fit in the pipeline buffers");

  const cnb_conv_params& p = a.p;
  const int M = a.M, K = a.K;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);   // [S]
  uint64_t* empty_bar = full_bar + S;                                       // [S]
  uint64_t* accum_bar = empty_bar + S;                                      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  long long* row_pix = reinterpret_cast<long long*>(smem + L::BAR_OFFSET + 256);   // [128] output pixel or -1
  int* row_b = reinterpret_cast<int*>(row_pix + BM);                                 // [128] batch index

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nkb = (K + BKE - 1) / BKE;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], NUM_PRODUCERS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < 4) {
    // ======================= producers: im2col gather (A) + weight rows (B) via cp.async =======================
    // Lane mapping is chunk-major: 8 consecutive lanes fetch the 8 16-byte chunks of ONE 128-byte row, so a warp
    // instruction touches 4 full cache lines (thread-per-row would touch 32 lines per instruction).
    const int c = tid & 7;                   // 16-byte chunk of the K slice owned by this thread
    const int rbase = tid >> 3;              // rows rbase + 16*j
    const int OHW = p.OH * p.OW;
    const char* in_base = reinterpret_cast<const char*>(p.in) + (size_t)p.in_coff * ELT;
    const char* w_base = HALF ? reinterpret_cast<const char*>(p.weight_lp) : reinterpret_cast<const char*>(p.weight);
    const size_t ldiB = (size_t)p.ldi * ELT;
    const uint32_t coff = (((uint32_t)c) ^ ((uint32_t)rbase & 7u)) << 4;   // swizzled chunk offset (row & 7 == rbase & 7)

    int a_pix[AROWS];          // ((b*H + iy0)*W + ix0) of the row's output pixel, or -1 when m >= M
    int a_yx[AROWS];           // iy0 | ix0 << 16
#pragma unroll
    for (int j = 0; j < AROWS; ++j) {
      const int m = m0 + rbase + 16 * j;
      if (m < M) {
        const int b = m / OHW;
        const int rem = m - b * OHW;
        const int oy = rem / p.OW;
        const int ox = rem - oy * p.OW;
        const int iy0 = oy * p.stride, ix0 = ox * p.stride;
        a_pix[j] = (b * p.H + iy0) * p.W + ix0;
        a_yx[j] = iy0 | (ix0 << 16);
      } else {
        a_pix[j] = -1;
        a_yx[j] = 0;
      }
    }

    const int total_it = nkb + S - 1;
    for (int it = 0; it < total_it; ++it) {
      if (it >= S - 1) {
        cp_async_wait<(S >= 2 ? S - 2 : 0)>();   // k-block (it - S + 1) has landed (S-2 younger groups may be in flight)
        fence_proxy_async();                     // generic-proxy smem writes -> visible to the tensor core
        mbar_arrive(&full_bar[(it - (S - 1)) % S]);
      }
      if (it < nkb) {
        const int s = it % S;
        if (it >= S) mbar_wait(&empty_bar[s], (uint32_t)(((it / S) - 1) & 1), abort_flag);
        const uint32_t st_base = smem_base + (uint32_t)s * L::STAGE_BYTES + coff;
        const int k = it * BKE + c * EPC;        // first K element of this thread's chunk
        const bool k_ok = k < K;
        int dy = 0, dx = 0, ch = 0;
        if (k_ok) {
          const int tap = k / p.Cin;
          ch = k - tap * p.Cin;
          dy = p.dy[tap];
          dx = p.dx[tap];
        }
        const int dpix = dy * p.W + dx;
        const char* src_c = in_base + (size_t)ch * ELT;
#pragma unroll
        for (int j = 0; j < AROWS; ++j) {
          const int iy = (a_yx[j] & 0xffff) + dy;
          const int ix = (a_yx[j] >> 16) + dx;
          const bool ok = k_ok && a_pix[j] >= 0 && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
          const char* src = ok ? src_c + (size_t)(a_pix[j] + dpix) * ldiB : in_base;
          cp_async16(st_base + (uint32_t)(rbase + 16 * j) * BKB, src, ok ? 16u : 0u);
        }
        const char* w_c = w_base + (size_t)k * ELT;
#pragma unroll
        for (int j = 0; j < BROWS; ++j) {
          const int br = rbase + 16 * j;
          if (BN % 16 == 0 || br < BN) {
            const bool ok = k_ok && (n0 + br) < p.Cout;
            const char* src = ok ? w_c + (size_t)(n0 + br) * K * ELT : w_base;
            cp_async16(st_base + A_STAGE_BYTES + (uint32_t)br * BKB, src, ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
    }

    // ======================= epilogue =======================================================================
    // TMEM -> registers (thread = accumulator row) -> smem staging (reusing the drained pipeline buffers) ->
    // coalesced channels-last store with the fused + bias + time-embedding row + residual (+SiLU).
    {
      const int m = m0 + tid;
      long long pix = -1;
      int b = 0;
      if (m < M) {
        b = m / OHW;
        const int rem = m - b * OHW;
        const int oy = rem / p.OW;
        const int ox = rem - oy * p.OW;
        pix = ((long long)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
      }
      row_pix[tid] = pix;
      row_b[tid] = b;
    }
    mbar_wait(accum_bar, 0, abort_flag);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem);
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)c0, v);
      float4* d = reinterpret_cast<float4*>(stage + (size_t)tid * OSTR + c0);
      d[0] = make_float4(v[0], v[1], v[2], v[3]);
      d[1] = make_float4(v[4], v[5], v[6], v[7]);
      d[2] = make_float4(v[8], v[9], v[10], v[11]);
      d[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");     // the 4 epilogue warps only
    if (!*abort_flag) {
      constexpr int C4 = BN / 4;                       // float4 columns per row
#pragma unroll 1
      for (int idx = tid; idx < BM * C4; idx += NUM_PRODUCERS) {
        const int row = idx / C4;
        const int c4 = idx - row * C4;
        const long long pix = row_pix[row];
        const int n = n0 + c4 * 4;
        if (pix >= 0 && n < p.Cout) {
          float4 o = *reinterpret_cast<const float4*>(stage + (size_t)row * OSTR + c4 * 4);
          if (p.bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          if (p.temb) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(
                p.temb + (size_t)(p.temb_per_sample ? row_b[row] : 0) * p.temb_ld + n));
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          if (p.residual) {
            float4 t;
            if (p.res_dtype == 1) {
              const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.residual) +
                                                                   pix * p.ldr + p.res_coff + n));
              const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
              const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
              t = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
              t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + pix * p.ldr +
                                                        p.res_coff + n));
            }
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          if (p.act == 1) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
          if (p.out_dtype == 1) {
            const __half2 lo = h2_sat(o.x, o.y), hi = h2_sat(o.z, o.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + pix * p.ldo + p.out_coff + n) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.ldo + p.out_coff + n) = o;
          }
        }
      }
    }
  } else {
    // ======================= MMA issuer (warp 4): one elected lane drives the tensor core ======================
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      const bool ok = mbar_wait(&full_bar[s], (uint32_t)((kb / S) & 1), abort_flag);
      tc_fence_after();
      if (lane == 0 && ok) {
        const uint32_t a_addr = smem_base + (uint32_t)s * L::STAGE_BYTES;
        const uint64_t adesc = make_desc_sw128(a_addr);
        const uint64_t bdesc = make_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BKB / 32; ++k) {
          // advance 32 bytes (8 tf32 / 16 fp16) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma<HALF>(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);              // frees the smem stage once these MMAs have read it
        if (kb == nkb - 1) umma_commit(accum_bar);   // accumulator complete -> epilogue
      }
      __syncwarp();
    }
    if (*abort_flag && lane == 0) mbar_arrive(accum_bar);   // drain path: never leave the epilogue waiting
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TMEM_COLS);
  }
}

template <int BN, int S, bool HALF>
static int launch_tc(const TcArgs& a, cudaStream_t st) {
  using L = SmemLayout<BN, S>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    CNB_CUDA(cudaFuncSetAttribute(conv_igemm_tc_kernel<BN, S, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  dim3 grid(ceil_div(a.M, BM), ceil_div(a.p.Cout, BN));
  conv_igemm_tc_kernel<BN, S, HALF><<<grid, 160, L::TOTAL, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

template <bool HALF>
static int dispatch_tc(const TcArgs& a, cudaStream_t st) {
  const int cout = a.p.Cout;
  static int deep = -1;
  if (deep < 0) {
    const char* e = getenv("CNB_TC_DEEP");
    deep = e ? atoi(e) : 1;
  }
  if (cout % 256 == 0) return launch_tc<256, 4, HALF>(a, st);
  if (cout % 128 == 0) return deep ? launch_tc<128, 6, HALF>(a, st) : launch_tc<128, 3, HALF>(a, st);
  if (cout % 64 == 0) return launch_tc<64, 4, HALF>(a, st);
  if (cout % 32 == 0) return launch_tc<32, 4, HALF>(a, st);
  return launch_tc<16, 4, HALF>(a, st);
}

}  // namespace tc

bool conv2d_tc_supported(const cnb_conv_params* p) {
  if (p->mode == CNB_MODE_F32) return false;
  const bool half = p->in_dtype == 1;
  const int epc = half ? 8 : 4;
  if (half && !p->weight_lp) return false;
  if (p->Cin % epc || p->ldi % epc || p->in_coff % epc) return false;
  if (p->Cout < 16 || p->Cout % 16) return false;
  if (p->ldo % 4 || p->out_coff % 4) return false;
  if (p->residual && (p->ldr % 4 || p->res_coff % 4)) return false;
  if (p->temb && (p->temb_ld % 4)) return false;
  if (((uintptr_t)p->in | (uintptr_t)p->weight | (uintptr_t)p->weight_lp | (uintptr_t)p->out | (uintptr_t)p->bias |
       (uintptr_t)p->temb | (uintptr_t)p->residual) & 15)
    return false;
  return true;
}

int conv2d_tc(const cnb_conv_params* p, cudaStream_t st) {
  tc::TcArgs a;
  a.p = *p;
  a.M = p->B * p->OH * p->OW;
  a.K = p->ntaps * p->Cin;
  return p->in_dtype == 1 ? tc::dispatch_tc<true>(a, st) : tc::dispatch_tc<false>(a, st);
}

}  // namespace cnb

extern "C" int cnb_tc_gemm_selftest(const float* a, const float* w, float* out, int M, int N, int K, int mode,
                                    cnb_stream_t stream) {
  cnb_conv_params p;
  memset(&p, 0, sizeof(p));
  p.in = a; p.weight = w; p.out = out;
  p.B = 1; p.H = M; p.W = 1; p.Cin = K; p.ldi = K; p.in_coff = 0;
  p.OH = M; p.OW = 1; p.OHf = M; p.OWf = 1;
  p.oy_mul = 1; p.ox_mul = 1;
  p.Cout = N; p.ldo = N; p.stride = 1; p.ntaps = 1;
  p.mode = mode;
  if (!cnb::conv2d_tc_supported(&p)) {
    cnb::set_error("tc_gemm_selftest: shape M=%d N=%d K=%d not supported by the tcgen05 path", M, N, K);
    return CNB_ERR_UNSUPPORTED;
  }
  return cnb::conv2d_tc(&p, (cudaStream_t)stream);
}

extern "C" int cnb_tc_error_flag(void) {
  int h = 0;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cnb::set_error("device error: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  const int h1 = cnb::tc_read_clear_error(), h2 = cnb::conv_tma_error_flag() | cnb::attention_tmem_error_flag();
  if (h1 < 0 || h2 < 0) {
    cnb::set_error("cudaMemcpyFromSymbol failed while reading the tcgen05 hang-guard flag");
    return CNB_ERR_CUDA;
  }
  h = h1 | h2;
  return h;
}
