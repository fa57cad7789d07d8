// Exact-fp32 flash-style self-attention core (CNB_MODE_F32) on packed projections.
//   qkv [B, L, 3E]  ->  out [B, L, E],   out_h = softmax(Q_h K_h^T / sqrt(d)) V_h,  d = E / heads.
// TPQ threads share one query row (each owns DT = d / TPQ channels of q and of the output accumulator);
// K/V tiles are staged in shared memory and read as warp broadcasts; the softmax is online (running max and
// sum), processed 8 keys at a time so the rescale is amortised.  Scores never touch global memory: at
// B = 1024, L = 784 a materialised fp32 score tensor would be 10 GB (hard part 5).
#include "common.cuh"

namespace cnb {

constexpr int ATT_THREADS = 128;
constexpr int CH = 8;  // keys per softmax chunk

template <int DT, int TPQ>
__global__ void __launch_bounds__(ATT_THREADS)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int E, int KT, float scale_log2) {
  constexpr int D = DT * TPQ;
  constexpr int QPB = ATT_THREADS / TPQ;
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;            // [KT][D]
  float* Vs = smem + KT * D;   // [KT][D]

  const int tid = threadIdx.x;
  const int part = tid % TPQ;
  const int ql = tid / TPQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q_idx = blockIdx.x * QPB + ql;
  const bool q_ok = q_idx < L;
  const size_t row3 = (size_t)3 * E;
  const float* base = qkv + (size_t)b * L * row3;

  float q[DT], acc[DT];
  {
    const float* qp = base + (size_t)(q_ok ? q_idx : 0) * row3 + h * D + part * DT;
#pragma unroll
    for (int i = 0; i < DT; i += 4) {
      float4 v = *reinterpret_cast<const float4*>(qp + i);
      q[i] = v.x * scale_log2; q[i + 1] = v.y * scale_log2; q[i + 2] = v.z * scale_log2; q[i + 3] = v.w * scale_log2;
    }
#pragma unroll
    for (int i = 0; i < DT; ++i) acc[i] = 0.f;
  }
  float m_run = -INFINITY, l_run = 0.f;

  for (int k0 = 0; k0 < L; k0 += KT) {
    const int kt = min(KT, L - k0);
    __syncthreads();
    // cooperative load of K and V rows [k0, k0+kt) for this head
    for (int u = tid; u < kt * (D / 4); u += ATT_THREADS) {
      int j = u / (D / 4), c4 = u % (D / 4);
      const float* kp = base + (size_t)(k0 + j) * row3 + E + h * D + c4 * 4;
      reinterpret_cast<float4*>(Ks + j * D)[c4] = *reinterpret_cast<const float4*>(kp);
      reinterpret_cast<float4*>(Vs + j * D)[c4] = *reinterpret_cast<const float4*>(kp + E);
    }
    __syncthreads();

    for (int j0 = 0; j0 < kt; j0 += CH) {
      float s[CH];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < CH; ++jj) {
        const int j = j0 + jj;
        float d = 0.f;
        if (j < kt) {
          const float* kr = Ks + j * D + part * DT;
#pragma unroll
          for (int i = 0; i < DT; i += 4) {
            float4 kv = *reinterpret_cast<const float4*>(kr + i);
            d = fmaf(q[i], kv.x, d); d = fmaf(q[i + 1], kv.y, d);
            d = fmaf(q[i + 2], kv.z, d); d = fmaf(q[i + 3], kv.w, d);
          }
        }
#pragma unroll
        for (int o = 1; o < TPQ; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s[jj] = (j < kt) ? d : -INFINITY;
        cmax = fmaxf(cmax, s[jj]);
      }
      const float m_new = fmaxf(m_run, cmax);
      const float corr = exp2f(m_run - m_new);   // m_run = -inf on the first chunk -> 0
      l_run *= corr;
#pragma unroll
      for (int i = 0; i < DT; ++i) acc[i] *= corr;
#pragma unroll
      for (int jj = 0; jj < CH; ++jj) {
        const int j = j0 + jj;
        if (j < kt) {
          const float pj = exp2f(s[jj] - m_new);
          l_run += pj;
          const float* vr = Vs + j * D + part * DT;
#pragma unroll
          for (int i = 0; i < DT; i += 4) {
            float4 vv = *reinterpret_cast<const float4*>(vr + i);
            acc[i] = fmaf(pj, vv.x, acc[i]); acc[i + 1] = fmaf(pj, vv.y, acc[i + 1]);
            acc[i + 2] = fmaf(pj, vv.z, acc[i + 2]); acc[i + 3] = fmaf(pj, vv.w, acc[i + 3]);
          }
        }
      }
      m_run = m_new;
    }
  }
  if (q_ok) {
    const float inv = 1.0f / l_run;
    float* op = out + ((size_t)b * L + q_idx) * E + h * D + part * DT;
#pragma unroll
    for (int i = 0; i < DT; i += 4)
      *reinterpret_cast<float4*>(op + i) = make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv);
  }
}

template <int DT, int TPQ>
static int launch_att(const float* qkv, float* out, int B, int L, int E, int heads, cudaStream_t st) {
  constexpr int D = DT * TPQ;
  constexpr int QPB = ATT_THREADS / TPQ;
  int KT = 6144 / D;            // K+V tiles <= 48 KB
  if (KT > 64) KT = 64;
  KT = (KT / CH) * CH;
  if (KT > ((L + CH - 1) / CH) * CH) KT = ((L + CH - 1) / CH) * CH;
  size_t smem = (size_t)2 * KT * D * sizeof(float);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  dim3 grid(ceil_div(L, QPB), heads, B);
  attention_f32_kernel<DT, TPQ><<<grid, ATT_THREADS, smem, st>>>(qkv, out, L, E, KT, scale_log2);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

int attention_f32(const float* qkv, float* out, int B, int L, int E, int heads, cudaStream_t st) {
  CNB_REQUIRE(heads > 0 && E % heads == 0, "attention: E=%d heads=%d", E, heads);
  CNB_REQUIRE(B <= 65535 && heads <= 65535, "attention: grid too large (B=%d)", B);
  const int d = E / heads;
  switch (d) {
    case 4: return launch_att<4, 1>(qkv, out, B, L, E, heads, st);
    case 8: return launch_att<8, 1>(qkv, out, B, L, E, heads, st);
    case 16: return launch_att<16, 1>(qkv, out, B, L, E, heads, st);
    case 24: return launch_att<24, 1>(qkv, out, B, L, E, heads, st);
    case 32: return launch_att<32, 1>(qkv, out, B, L, E, heads, st);
    case 48: return launch_att<48, 1>(qkv, out, B, L, E, heads, st);
    case 64: return launch_att<64, 1>(qkv, out, B, L, E, heads, st);
    case 96: return launch_att<48, 2>(qkv, out, B, L, E, heads, st);
    case 128: return launch_att<64, 2>(qkv, out, B, L, E, heads, st);
    case 192: return launch_att<48, 4>(qkv, out, B, L, E, heads, st);
    default:
      set_error("attention: head dim %d not instantiated", d);
      return CNB_ERR_UNSUPPORTED;
  }
}

}  // namespace cnb
