// GroupNorm (+SiLU) over channels-last fp32 activations [B, HW, C]  (memory-bound: judged by HBM GB/s).
// One CTA per sample: the [HW, C] slab is contiguous, so every global access is a fully coalesced float4.
// Statistics are two-pass (mean, then centred sum of squares) to match ATen's numerics; the second and third
// sweeps of the slab are L1/L2 hits (a slab is at most a few hundred KB), so DRAM sees one read + one write.
// Per-thread column ownership (thread <-> fixed 4 channels) makes the reduction a deterministic tree:
// registers -> smem [rows][C] -> per-channel -> per-group.
#include <cuda_fp16.h>

#include "common.cuh"

namespace cnb {

template <bool HALF> struct OutVec;
template <> struct OutVec<false> {
  using type = float4;
  static __device__ __forceinline__ float4 pack(float4 v) { return v; }
};
template <> struct OutVec<true> {
  using type = uint2;   // 4 x fp16
  static __device__ __forceinline__ uint2 pack(float4 v) {
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
};

template <bool SILU, bool HALF>
__global__ void __launch_bounds__(1024)
groupnorm_nhwc_kernel(const float* __restrict__ x, void* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int HW, int C, int G, float eps) {
  extern __shared__ float sm[];
  const int C4 = C >> 2;
  const int NT = blockDim.x;
  const int R = NT / C4;            // pixel rows swept per iteration
  float* part = sm;                 // [R][C]
  float* chan = part + R * C;       // [C]
  float* g_mean = chan + C;         // [G]
  float* g_rstd = g_mean + G;       // [G]

  const int tid = threadIdx.x;
  const int col = tid % C4;
  const int r0 = tid / C4;
  const int cg = C / G;
  const float4* xs = reinterpret_cast<const float4*>(x + (size_t)blockIdx.x * HW * C);
  using OV = OutVec<HALF>;
  typename OV::type* ys = reinterpret_cast<typename OV::type*>(y) + (size_t)blockIdx.x * HW * C4;
  const float inv_n = 1.0f / (float)(cg * HW);

  // ---- pass 1: per-channel sums -> group means
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int pidx = r0;
  for (; pidx + 3 * R < HW; pidx += 4 * R) {
    float4 v0 = xs[(size_t)pidx * C4 + col];
    float4 v1 = xs[(size_t)(pidx + R) * C4 + col];
    float4 v2 = xs[(size_t)(pidx + 2 * R) * C4 + col];
    float4 v3 = xs[(size_t)(pidx + 3 * R) * C4 + col];
    s.x += (v0.x + v1.x) + (v2.x + v3.x);
    s.y += (v0.y + v1.y) + (v2.y + v3.y);
    s.z += (v0.z + v1.z) + (v2.z + v3.z);
    s.w += (v0.w + v1.w) + (v2.w + v3.w);
  }
  for (; pidx < HW; pidx += R) {
    float4 v = xs[(size_t)pidx * C4 + col];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  reinterpret_cast<float4*>(part + r0 * C)[col] = s;
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += part[r * C + c];
    chan[c] = t;
  }
  __syncthreads();
  for (int g = tid; g < G; g += NT) {
    float t = 0.f;
    for (int c = 0; c < cg; ++c) t += chan[g * cg + c];
    g_mean[g] = t * inv_n;
  }
  __syncthreads();

  // ---- pass 2: centred second moment -> rstd
  const int c0 = col * 4;
  const float m0 = g_mean[(c0 + 0) / cg], m1 = g_mean[(c0 + 1) / cg];
  const float m2 = g_mean[(c0 + 2) / cg], m3 = g_mean[(c0 + 3) / cg];
  s = make_float4(0.f, 0.f, 0.f, 0.f);
  pidx = r0;
  for (; pidx + 3 * R < HW; pidx += 4 * R) {
    float4 v0 = xs[(size_t)pidx * C4 + col];
    float4 v1 = xs[(size_t)(pidx + R) * C4 + col];
    float4 v2 = xs[(size_t)(pidx + 2 * R) * C4 + col];
    float4 v3 = xs[(size_t)(pidx + 3 * R) * C4 + col];
    float d;
    d = v0.x - m0; s.x = fmaf(d, d, s.x); d = v1.x - m0; s.x = fmaf(d, d, s.x);
    d = v2.x - m0; s.x = fmaf(d, d, s.x); d = v3.x - m0; s.x = fmaf(d, d, s.x);
    d = v0.y - m1; s.y = fmaf(d, d, s.y); d = v1.y - m1; s.y = fmaf(d, d, s.y);
    d = v2.y - m1; s.y = fmaf(d, d, s.y); d = v3.y - m1; s.y = fmaf(d, d, s.y);
    d = v0.z - m2; s.z = fmaf(d, d, s.z); d = v1.z - m2; s.z = fmaf(d, d, s.z);
    d = v2.z - m2; s.z = fmaf(d, d, s.z); d = v3.z - m2; s.z = fmaf(d, d, s.z);
    d = v0.w - m3; s.w = fmaf(d, d, s.w); d = v1.w - m3; s.w = fmaf(d, d, s.w);
    d = v2.w - m3; s.w = fmaf(d, d, s.w); d = v3.w - m3; s.w = fmaf(d, d, s.w);
  }
  for (; pidx < HW; pidx += R) {
    float4 v = xs[(size_t)pidx * C4 + col];
    float d;
    d = v.x - m0; s.x = fmaf(d, d, s.x);
    d = v.y - m1; s.y = fmaf(d, d, s.y);
    d = v.z - m2; s.z = fmaf(d, d, s.z);
    d = v.w - m3; s.w = fmaf(d, d, s.w);
  }
  reinterpret_cast<float4*>(part + r0 * C)[col] = s;
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += part[r * C + c];
    chan[c] = t;
  }
  __syncthreads();
  for (int g = tid; g < G; g += NT) {
    float t = 0.f;
    for (int c = 0; c < cg; ++c) t += chan[g * cg + c];
    g_rstd[g] = rsqrtf(t * inv_n + eps);
  }
  __syncthreads();

  // ---- pass 3: normalise, affine, (SiLU), store
  const float4 ga = reinterpret_cast<const float4*>(gamma)[col];
  const float4 be = reinterpret_cast<const float4*>(beta)[col];
  const float a0 = g_rstd[(c0 + 0) / cg] * ga.x, a1 = g_rstd[(c0 + 1) / cg] * ga.y;
  const float a2 = g_rstd[(c0 + 2) / cg] * ga.z, a3 = g_rstd[(c0 + 3) / cg] * ga.w;
  const float b0 = be.x - m0 * a0, b1 = be.y - m1 * a1, b2 = be.z - m2 * a2, b3 = be.w - m3 * a3;
  auto apply = [&](float4 v) {
    float4 o;
    o.x = fmaf(v.x, a0, b0); o.y = fmaf(v.y, a1, b1); o.z = fmaf(v.z, a2, b2); o.w = fmaf(v.w, a3, b3);
    if (SILU) { o.x = silu_f(o.x); o.y = silu_f(o.y); o.z = silu_f(o.z); o.w = silu_f(o.w); }
    return o;
  };
  pidx = r0;
  for (; pidx + 3 * R < HW; pidx += 4 * R) {
    float4 v0 = xs[(size_t)pidx * C4 + col];
    float4 v1 = xs[(size_t)(pidx + R) * C4 + col];
    float4 v2 = xs[(size_t)(pidx + 2 * R) * C4 + col];
    float4 v3 = xs[(size_t)(pidx + 3 * R) * C4 + col];
    ys[(size_t)pidx * C4 + col] = OV::pack(apply(v0));
    ys[(size_t)(pidx + R) * C4 + col] = OV::pack(apply(v1));
    ys[(size_t)(pidx + 2 * R) * C4 + col] = OV::pack(apply(v2));
    ys[(size_t)(pidx + 3 * R) * C4 + col] = OV::pack(apply(v3));
  }
  for (; pidx < HW; pidx += R) ys[(size_t)pidx * C4 + col] = OV::pack(apply(xs[(size_t)pidx * C4 + col]));
}

int groupnorm(const float* x, void* y, const float* gamma, const float* beta, int B, int HW, int C, int G,
              float eps, int silu, int out_f16, cudaStream_t st) {
  CNB_REQUIRE(C % 4 == 0 && C % G == 0 && C / 4 <= 1024, "groupnorm: C=%d G=%d unsupported", C, G);
  const int C4 = C / 4;
  int R = 512 / C4;
  if (R < 1) R = 1;
  if (R > HW) R = HW;
  const int NT = R * C4;
  size_t smem = ((size_t)R * C + C + 2 * G) * sizeof(float);
  CNB_REQUIRE(smem <= 48 * 1024, "groupnorm: smem %zu too large", smem);
  if (silu && out_f16)
    groupnorm_nhwc_kernel<true, true><<<B, NT, smem, st>>>(x, y, gamma, beta, HW, C, G, eps);
  else if (silu)
    groupnorm_nhwc_kernel<true, false><<<B, NT, smem, st>>>(x, y, gamma, beta, HW, C, G, eps);
  else if (out_f16)
    groupnorm_nhwc_kernel<false, true><<<B, NT, smem, st>>>(x, y, gamma, beta, HW, C, G, eps);
  else
    groupnorm_nhwc_kernel<false, false><<<B, NT, smem, st>>>(x, y, gamma, beta, HW, C, G, eps);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace cnb
