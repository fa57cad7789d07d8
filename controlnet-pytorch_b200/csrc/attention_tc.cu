// Tensor-core flash attention for the tf32 mode:  qkv [B, L, 3E] -> out [B, L, E].
//
// Head dims on this path are tiny (4..64; 128 at CIFAR's 8x8 level) and L <= 1024, so the kernel is bound by the
// softmax (MUFU ex2 + FP32 pipe), not by the tensor pipe: per score it does 2*2*d tensor FLOPs but one exp2 and ~6
// FP32 ops.  A 128-lane tcgen05/TMEM tile would need head dims padded to 32 columns and a TMEM round trip for P; the
// register-resident FA2 formulation below keeps S and P in the MMA fragment registers instead:
//   * one CTA = one (batch, head) x 64 queries, 4 warps x 16 query rows, K/V tiles of BKEYS keys double-buffered
//     in shared memory with 16-byte cp.async (row stride D+4 floats -> conflict-free fragment reads);
//   * S = Q K^T and O += P V use mma.sync.m16n8k8 tf32 with fp32 accumulate; the S accumulator fragment is reused
//     directly as the A fragment of the second MMA by relabelling the k index (hardware k = lane%4 <-> key
//     2*(lane%4), k = lane%4 + 4 <-> key 2*(lane%4)+1) and fetching V rows in the same order;
//   * online softmax in the exp2 domain (scale * log2(e) folded into Q), row max / sum reduced over the 4 lanes of
//     a quad with shuffles; scores never leave registers.
#include "common.cuh"

namespace cnb {

namespace atc {

constexpr int BQ = 64;        // queries per CTA
constexpr int THREADS = 128;  // 4 warps

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <int D, int BKEYS>
__global__ void __launch_bounds__(THREADS)
attention_tf32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int E, int d_real,
                      float scale_log2) {
  constexpr int STRIDE = D + 4;
  constexpr int NT = BKEYS / 8;    // S n-tiles (8 keys each) == k-steps of the PV product
  constexpr int KS = D / 8;        // k-steps of QK^T == n-tiles of O
  constexpr int TILE = BKEYS * STRIDE;
  extern __shared__ __align__(16) float smem[];   // [2 stages][K | V][BKEYS][STRIDE]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * BQ + warp * 16;
  const size_t row3 = (size_t)3 * E;
  const float* base = qkv + (size_t)b * L * row3 + (size_t)h * d_real;
  const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem);

  // ---- Q fragments (scaled), kept in registers for the whole kernel
  uint32_t qf[KS][4];
  {
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c0 = ks * 8 + q4, c1 = c0 + 4;
      float v00 = 0.f, v10 = 0.f, v01 = 0.f, v11 = 0.f;
      if (r0 < L) {
        if (c0 < d_real) v00 = base[(size_t)r0 * row3 + c0];
        if (c1 < d_real) v01 = base[(size_t)r0 * row3 + c1];
      }
      if (r1 < L) {
        if (c0 < d_real) v10 = base[(size_t)r1 * row3 + c0];
        if (c1 < d_real) v11 = base[(size_t)r1 * row3 + c1];
      }
      qf[ks][0] = __float_as_uint(v00 * scale_log2);
      qf[ks][1] = __float_as_uint(v10 * scale_log2);
      qf[ks][2] = __float_as_uint(v01 * scale_log2);
      qf[ks][3] = __float_as_uint(v11 * scale_log2);
    }
  }

  float o[KS][4];
#pragma unroll
  for (int i = 0; i < KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = (L + BKEYS - 1) / BKEYS;
  constexpr int CHUNKS = D / 4;                    // 16-byte chunks per K (or V) row
  auto load_tile = [&](int t, int stage) {
    const int k0 = t * BKEYS;
    const uint32_t sK = smem_u + (uint32_t)(stage * 2 * TILE) * 4u;
    const uint32_t sV = sK + (uint32_t)TILE * 4u;
    for (int u = tid; u < BKEYS * CHUNKS; u += THREADS) {
      const int j = u / CHUNKS, c = u - j * CHUNKS;
      const int key = k0 + j;
      const bool ok = (key < L) && (c * 4 < d_real);
      const float* kp = base + (size_t)(ok ? key : 0) * row3 + E + (ok ? c * 4 : 0);
      const uint32_t off = (uint32_t)(j * STRIDE + c * 4) * 4u;
      cp16(sK + off, kp, ok ? 16u : 0u);
      cp16(sV + off, kp + E, ok ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  load_tile(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int stage = t & 1;
    if (t + 1 < ntiles) {
      load_tile(t + 1, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* Ks = smem + stage * 2 * TILE;
    const float* Vs = Ks + TILE;

    // ---- S = Q K^T  (16 x BKEYS per warp)
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const float* kr = Ks + (nt * 8 + g) * STRIDE + q4;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t b0 = __float_as_uint(kr[ks * 8]);
        const uint32_t b1 = __float_as_uint(kr[ks * 8 + 4]);
        mma_tf32(s[nt], qf[ks], b0, b1);
      }
    }
    // ---- mask keys beyond L (last tile only)
    const int kbase = t * BKEYS;
    if (kbase + BKEYS > L) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int key = kbase + nt * 8 + 2 * q4;
        if (key >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (key + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g and g+8)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile has >= 1 valid key
    const float cr0 = exp2f(m0 - mn0), cr1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= cr0; l1 *= cr1;
#pragma unroll
    for (int i = 0; i < KS; ++i) { o[i][0] *= cr0; o[i][1] *= cr0; o[i][2] *= cr1; o[i][3] *= cr1; }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    // ---- O += P V : the S fragment is the A fragment (k relabelled: a0/a1 <- key 2*q4, a2/a3 <- key 2*q4+1)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t pa[4] = {__float_as_uint(s[nt][0]), __float_as_uint(s[nt][2]), __float_as_uint(s[nt][1]),
                        __float_as_uint(s[nt][3])};
      const float* vr = Vs + (nt * 8 + 2 * q4) * STRIDE + g;
#pragma unroll
      for (int i = 0; i < KS; ++i) {
        const uint32_t b0 = __float_as_uint(vr[i * 8]);
        const uint32_t b1 = __float_as_uint(vr[STRIDE + i * 8]);
        mma_tf32(o[i], pa, b0, b1);
      }
    }
    __syncthreads();   // all warps done with this stage before it is refilled
  }

  // ---- normalise and store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  float* ob = out + (size_t)b * L * E + (size_t)h * d_real;
#pragma unroll
  for (int i = 0; i < KS; ++i) {
    const int c = i * 8 + 2 * q4;
    if (c < d_real) {   // d_real is even, so c + 1 < d_real as well
      if (r0 < L) *reinterpret_cast<float2*>(ob + (size_t)r0 * E + c) = make_float2(o[i][0] * i0, o[i][1] * i0);
      if (r1 < L) *reinterpret_cast<float2*>(ob + (size_t)r1 * E + c) = make_float2(o[i][2] * i1, o[i][3] * i1);
    }
  }
}

template <int D, int BKEYS>
static int launch(const float* qkv, float* out, int B, int L, int E, int heads, int d_real, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)2 * 2 * BKEYS * (D + 4) * sizeof(float);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    CNB_CUDA(cudaFuncSetAttribute(attention_tf32_kernel<D, BKEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)d_real);
  dim3 grid(ceil_div(L, BQ), heads, B);
  attention_tf32_kernel<D, BKEYS><<<grid, THREADS, SMEM, st>>>(qkv, out, L, E, d_real, scale_log2);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace atc

bool attention_tc_supported(int E, int heads) {
  if (heads <= 0 || E % heads) return false;
  const int d = E / heads;
  return d == 4 || d == 8 || d == 16 || d == 24 || d == 32 || d == 48 || d == 64 || d == 128;
}

int attention_tc(const float* qkv, float* out, int B, int L, int E, int heads, cudaStream_t st) {
  CNB_REQUIRE(B <= 65535 && heads <= 65535, "attention: grid too large (B=%d)", B);
  const int d = E / heads;
  switch (d) {
    case 4: return atc::launch<8, 64>(qkv, out, B, L, E, heads, d, st);
    case 8: return atc::launch<8, 64>(qkv, out, B, L, E, heads, d, st);
    case 16: return atc::launch<16, 64>(qkv, out, B, L, E, heads, d, st);
    case 24: return atc::launch<24, 64>(qkv, out, B, L, E, heads, d, st);
    case 32: return atc::launch<32, 64>(qkv, out, B, L, E, heads, d, st);
    case 48: return atc::launch<48, 64>(qkv, out, B, L, E, heads, d, st);
    case 64: return atc::launch<64, 64>(qkv, out, B, L, E, heads, d, st);
    case 128: return atc::launch<128, 32>(qkv, out, B, L, E, heads, d, st);
    default:
      set_error("attention_tc: head dim %d not instantiated", d);
      return CNB_ERR_UNSUPPORTED;
  }
}

}  // namespace cnb
