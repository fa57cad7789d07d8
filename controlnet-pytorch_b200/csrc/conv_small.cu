// Direct CUDA-core convolutions for the tiny-channel ends of the U-Nets (HBM-bound; judged by GB/s, exact fp32 FFMA):
//   conv_small_cin  : Cin <= 4  (conv_in with 1/3/4 image channels, first hint conv with 3 channels), Cout % 4 == 0.
//                     A lane owns 4 output channels (weights [K][Cout] in smem, float4 per k); the Cout/4 lanes of
//                     a pixel read the same input taps (L1 broadcast) and together write one contiguous
//                     channels-last row segment, so a warp store is a run of full 128-byte lines.
//   conv_small_cout : Cout <= 4 (conv_out to 1/3/4 image channels), Cin % 4 == 0.  One thread per output pixel; the
//                     3x3 neighbourhood is re-read from L1 by the neighbouring threads, weights are warp-uniform
//                     shared-memory broadcasts.
// Both fuse + bias + time-embedding row + residual (+SiLU) like the tensor-core kernels.
#include <cuda_fp16.h>

#include "common.cuh"

namespace cnb {

struct SmallArgs {
  cnb_conv_params p;
  int M, K;
};

__global__ void __launch_bounds__(256)
conv_small_cin_kernel(const __grid_constant__ SmallArgs a) {
  const cnb_conv_params& p = a.p;
  const int K = a.K;                       // ntaps * Cin
  const int lpr = p.Cout >> 2;             // lanes per pixel
  const int ppb = blockDim.x / lpr;        // pixels per CTA iteration
  const int lane_c = threadIdx.x % lpr;
  const int lane_p = threadIdx.x / lpr;
  const int n = lane_c * 4;
  const float* in = reinterpret_cast<const float*>(p.in);
  extern __shared__ float wsm[];           // [K][Cout]: a lane reads the float4 of its 4 channels, pixels broadcast
  for (int i = threadIdx.x; i < K * p.Cout; i += blockDim.x) {
    const int k = i / p.Cout, o = i - k * p.Cout;
    wsm[i] = __ldg(p.weight + (size_t)o * K + k);
  }
  __syncthreads();
  const float4* w4 = reinterpret_cast<const float4*>(wsm) + lane_c;

  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  if (p.temb && !p.temb_per_sample) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p.temb + n));
    bv.x += t.x; bv.y += t.y; bv.z += t.z; bv.w += t.w;
  }
  const int OHW = p.OH * p.OW;
  const int Cin = p.Cin;
  for (int m = blockIdx.x * ppb + lane_p; m < a.M; m += gridDim.x * ppb) {
    const int b = m / OHW;
    const int rem = m - b * OHW;
    const int oy = rem / p.OW;
    const int ox = rem - oy * p.OW;
    const int iy0 = oy * p.stride, ix0 = ox * p.stride;
    float4 acc = bv;
    for (int tap = 0; tap < p.ntaps; ++tap) {
      const int iy = iy0 + p.dy[tap], ix = ix0 + p.dx[tap];
      if ((unsigned)iy >= (unsigned)p.H || (unsigned)ix >= (unsigned)p.W) continue;
      const float* src = in + ((size_t)(b * p.H + iy) * p.W + ix) * p.ldi + p.in_coff;
      const float4* wt = w4 + (size_t)tap * Cin * lpr;
      for (int c = 0; c < Cin; ++c) {
        const float x = __ldg(src + c);
        const float4 wk = wt[(size_t)c * lpr];
        acc.x = fmaf(x, wk.x, acc.x); acc.y = fmaf(x, wk.y, acc.y);
        acc.z = fmaf(x, wk.z, acc.z); acc.w = fmaf(x, wk.w, acc.w);
      }
    }
    const size_t pix = ((size_t)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
    if (p.temb && p.temb_per_sample) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.temb + (size_t)b * p.temb_ld + n));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
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
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (p.act == 1) { acc.x = silu_f(acc.x); acc.y = silu_f(acc.y); acc.z = silu_f(acc.z); acc.w = silu_f(acc.w); }
    if (p.out_dtype == 1) {
      const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi = __floats2half2_rn(acc.z, acc.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + pix * p.ldo + p.out_coff + n) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.ldo + p.out_coff + n) = acc;
    }
  }
}

template <int COUT, bool HALF_IN>
__global__ void __launch_bounds__(128)
conv_small_cout_kernel(const __grid_constant__ SmallArgs a) {
  const cnb_conv_params& p = a.p;
  const int K = a.K;
  extern __shared__ float wsm[];           // [COUT][K]
  for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) wsm[i] = __ldg(p.weight + i);
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const int OHW = p.OH * p.OW;
  const int b = m / OHW;
  const int rem = m - b * OHW;
  const int oy = rem / p.OW;
  const int ox = rem - oy * p.OW;
  const int iy0 = oy * p.stride, ix0 = ox * p.stride;
  const int C4 = p.Cin >> 2;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int iy = iy0 + p.dy[tap], ix = ix0 + p.dx[tap];
    if ((unsigned)iy >= (unsigned)p.H || (unsigned)ix >= (unsigned)p.W) continue;
    const size_t off = ((size_t)(b * p.H + iy) * p.W + ix) * p.ldi + p.in_coff;
    const float* wt = wsm + tap * p.Cin;
#pragma unroll 4
    for (int c4 = 0; c4 < C4; ++c4) {
      float4 x;
      if (HALF_IN) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.in) + off) + c4);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        x = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        x = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.in) + off) + c4);
      }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float4 wv = *reinterpret_cast<const float4*>(wt + (size_t)o * K + c4 * 4);
        acc[o] = fmaf(x.x, wv.x, acc[o]); acc[o] = fmaf(x.y, wv.y, acc[o]);
        acc[o] = fmaf(x.z, wv.z, acc[o]); acc[o] = fmaf(x.w, wv.w, acc[o]);
      }
    }
  }
  const size_t pix = ((size_t)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float v = acc[o];
    if (p.bias) v += __ldg(p.bias + o);
    if (p.temb) v += __ldg(p.temb + (size_t)(p.temb_per_sample ? b : 0) * p.temb_ld + o);
    if (p.residual)
      v += p.res_dtype == 1 ? __half2float(reinterpret_cast<const __half*>(p.residual)[pix * p.ldr + p.res_coff + o])
                            : __ldg(reinterpret_cast<const float*>(p.residual) + pix * p.ldr + p.res_coff + o);
    if (p.act == 1) v = silu_f(v);
    reinterpret_cast<float*>(p.out)[pix * p.ldo + p.out_coff + o] = v;
  }
}

static bool small_cin_ok(const cnb_conv_params* p) {
  const int K = p->ntaps * p->Cin;
  return p->Cin <= 4 && p->in_dtype == 0 && p->Cout % 4 == 0 && p->Cout >= 4 && p->Cout <= 256 && K <= 64 &&
         256 % (p->Cout / 4) == 0 && p->ldo % 4 == 0 && p->out_coff % 4 == 0 &&
         (!p->residual || (p->ldr % 4 == 0 && p->res_coff % 4 == 0)) && (!p->temb || p->temb_ld % 4 == 0) &&
         (((uintptr_t)p->out | (uintptr_t)p->bias | (uintptr_t)p->temb | (uintptr_t)p->residual) & 15) == 0;
}

static bool small_cout_ok(const cnb_conv_params* p) {
  const int K = p->ntaps * p->Cin;
  return p->Cout <= 4 && p->Cin % 4 == 0 && p->out_dtype == 0 && p->ldi % 4 == 0 && p->in_coff % 4 == 0 &&
         (size_t)p->Cout * K * sizeof(float) <= 48 * 1024 && ((uintptr_t)p->in & 15) == 0;
}

bool conv2d_small_supported(const cnb_conv_params* p) { return small_cin_ok(p) || small_cout_ok(p); }

int conv2d_small(const cnb_conv_params* p, cudaStream_t st) {
  SmallArgs a;
  a.p = *p;
  a.M = p->B * p->OH * p->OW;
  a.K = p->ntaps * p->Cin;
  if (small_cin_ok(p)) {
    const int ppb = 256 / (p->Cout / 4);
    int grid = ceil_div(a.M, ppb);
    if (grid > 148 * 8) grid = 148 * 8;           // grid-stride: the weight matrix is staged in smem once per CTA
    conv_small_cin_kernel<<<grid, 256, (size_t)a.K * p->Cout * sizeof(float), st>>>(a);
    CNB_LAUNCH_CHECK();
    return CNB_OK;
  }
  const size_t smem = (size_t)p->Cout * a.K * sizeof(float);
  const int grid = ceil_div(a.M, 128);
  const bool h = p->in_dtype == 1;
  switch (p->Cout) {
    case 1: h ? conv_small_cout_kernel<1, true><<<grid, 128, smem, st>>>(a) : conv_small_cout_kernel<1, false><<<grid, 128, smem, st>>>(a); break;
    case 2: h ? conv_small_cout_kernel<2, true><<<grid, 128, smem, st>>>(a) : conv_small_cout_kernel<2, false><<<grid, 128, smem, st>>>(a); break;
    case 3: h ? conv_small_cout_kernel<3, true><<<grid, 128, smem, st>>>(a) : conv_small_cout_kernel<3, false><<<grid, 128, smem, st>>>(a); break;
    default: h ? conv_small_cout_kernel<4, true><<<grid, 128, smem, st>>>(a) : conv_small_cout_kernel<4, false><<<grid, 128, smem, st>>>(a); break;
  }
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace cnb
