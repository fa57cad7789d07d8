// tcgen05 implicit-GEMM convolution (CNB_MODE_TF32): the tensor-core path of cnb_conv2d on sm_100a.
//
//   GEMM view        M = B*OH*OW output pixels (tile 128 = one UMMA M),  N = Cout (tile BN in {16,32,64,128}),
//                    K = ntaps*Cin (tile 32 fp32 = one 128-byte swizzle row).
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
#include "common.cuh"

namespace cnb {

__device__ int g_tc_error = 0;

namespace tc {

constexpr int BM = 128;
constexpr int BKB = 128;                 // bytes of K per stage row (one SWIZZLE_128B row)
constexpr int A_STAGE_BYTES = BM * BKB;  // 16 KB
constexpr int NUM_PRODUCERS = 128;
constexpr long long WAIT_LIMIT_CYCLES = 4000000000ll;   // ~2 s at 1.9 GHz

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
// Bounded wait.  Returns false (and latches *abort) if the phase did not complete within the guard.
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
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (BF16) {
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
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool bf16) {
  uint32_t fmt = bf16 ? 1u : 2u;   // F16F32Format: BF16 = 1, TF32 = 2
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
  cnb_conv_params p;
  int M, K;
};

template <int BN, int S>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BN * BKB;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = S * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers/tmem slot + 1 KB alignment slack
};

template <int BN, int S>
__global__ void __launch_bounds__(160, 1)
conv_igemm_tf32_kernel(const __grid_constant__ TcArgs a) {
  using L = SmemLayout<BN, S>;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = make_idesc(BM, BN, false);

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

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nkb = (K * 4 + BKB - 1) / BKB;

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
    const int r = tid;                       // tile row owned by this thread (A: pixel m0+r, B: channel n0+r)
    const int m = m0 + r;
    const bool m_ok = m < M;
    const int OHW = p.OH * p.OW;
    int b = 0, oy = 0, ox = 0;
    if (m_ok) {
      b = m / OHW;
      int rem = m - b * OHW;
      oy = rem / p.OW;
      ox = rem - oy * p.OW;
    }
    const int iy0 = oy * p.stride, ix0 = ox * p.stride;
    const float* in_b = p.in + (size_t)b * p.H * p.W * p.ldi + p.in_coff;
    const bool b_row = (r < BN);
    const bool n_ok = b_row && (n0 + r < p.Cout);
    const float* w_row = p.weight + (size_t)(n_ok ? (n0 + r) : 0) * K;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t a_row_off = (uint32_t)r * BKB;
    const uint32_t b_row_off = A_STAGE_BYTES + (uint32_t)r * BKB;

    const int total_it = nkb + S - 1;
    for (int it = 0; it < total_it; ++it) {
      if (it < nkb) {
        const int s = it % S;
        if (it >= S) mbar_wait(&empty_bar[s], (uint32_t)(((it / S) - 1) & 1), abort_flag);
        const uint32_t st_base = smem_base + (uint32_t)s * L::STAGE_BYTES;
        const int k0 = it * (BKB / 4);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int k = k0 + c * 4;
          const float* src = p.in;
          uint32_t nbytes = 0;
          if (m_ok && k < K) {
            const int tap = k / p.Cin;
            const int ch = k - tap * p.Cin;
            const int iy = iy0 + p.dy[tap];
            const int ix = ix0 + p.dx[tap];
            if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) {
              src = in_b + ((size_t)iy * p.W + ix) * p.ldi + ch;
              nbytes = 16;
            }
          }
          cp_async16(st_base + a_row_off + (((uint32_t)c ^ swz) << 4), src, nbytes);
        }
        if (b_row) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k = k0 + c * 4;
            const bool ok = n_ok && (k < K);
            cp_async16(st_base + b_row_off + (((uint32_t)c ^ swz) << 4), ok ? (const void*)(w_row + k) : (const void*)p.weight,
                       ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      if (it >= S - 1) {
        cp_async_wait<S - 1>();       // k-block (it - S + 1) has landed
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&full_bar[(it - (S - 1)) % S]);
      }
    }

    // ======================= epilogue: TMEM -> registers -> fused adds -> channels-last store ===================
    mbar_wait(accum_bar, 0, abort_flag);
    tc_fence_after();
    long long pix = 0;
    if (m_ok) pix = ((long long)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
    float* dst = p.out + pix * p.ldo + p.out_coff;
    const float* res = p.residual ? p.residual + pix * p.ldr + p.res_coff : nullptr;
    const float* te = p.temb ? p.temb + (size_t)(p.temb_per_sample ? b : 0) * p.temb_ld : nullptr;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(lane_addr + (uint32_t)c0, v);
      if (m_ok && !*abort_flag) {
        const int nb = n0 + c0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (nb + j < p.Cout) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.bias) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
              o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (te) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(te + nb + j));
              o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (res) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(res + nb + j));
              o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (p.act == 1) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
            *reinterpret_cast<float4*>(dst + nb + j) = o;
          }
        }
      }
    }
    tc_fence_before();
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
          // advance 32 bytes (8 tf32) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma<false>(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (kb | k) ? 1u : 0u);
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

template <int BN, int S>
static int launch_tc(const TcArgs& a, cudaStream_t st) {
  using L = SmemLayout<BN, S>;
  static bool attr_set = false;
  if (!attr_set) {
    CNB_CUDA(cudaFuncSetAttribute(conv_igemm_tf32_kernel<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.M, BM), ceil_div(a.p.Cout, BN));
  conv_igemm_tf32_kernel<BN, S><<<grid, 160, L::TOTAL, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace tc

bool conv2d_tc_supported(const cnb_conv_params* p) {
  if (p->mode != CNB_MODE_TF32) return false;   // bf16 operands: next round (activations are fp32 in HBM today)
  if (p->Cin % 4 || p->ldi % 4 || p->in_coff % 4) return false;
  if (p->Cout < 16 || p->Cout % 16) return false;
  if (p->ldo % 4 || p->out_coff % 4) return false;
  if (p->residual && (p->ldr % 4 || p->res_coff % 4)) return false;
  if (p->temb && (p->temb_ld % 4)) return false;
  if (((uintptr_t)p->in | (uintptr_t)p->weight | (uintptr_t)p->out | (uintptr_t)p->bias | (uintptr_t)p->temb |
       (uintptr_t)p->residual) & 15)
    return false;
  return true;
}

int conv2d_tc(const cnb_conv_params* p, cudaStream_t st) {
  tc::TcArgs a;
  a.p = *p;
  a.M = p->B * p->OH * p->OW;
  a.K = p->ntaps * p->Cin;
  if (p->Cout % 128 == 0) return tc::launch_tc<128, 3>(a, st);
  if (p->Cout % 64 == 0) return tc::launch_tc<64, 4>(a, st);
  if (p->Cout % 32 == 0) return tc::launch_tc<32, 4>(a, st);
  return tc::launch_tc<16, 4>(a, st);
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
  e = cudaMemcpyFromSymbol(&h, cnb::g_tc_error, sizeof(int));
  if (e != cudaSuccess) {
    cnb::set_error("cudaMemcpyFromSymbol: %s", cudaGetErrorString(e));
    return CNB_ERR_CUDA;
  }
  if (h) {
    int z = 0;
    cudaMemcpyToSymbol(cnb::g_tc_error, &z, sizeof(int));
  }
  return h;
}
