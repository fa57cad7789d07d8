// Direct CUDA-core convolutions for the tiny-channel ends of the U-Nets (HBM-bound; judged by GB/s, exact fp32 FFMA):
//   conv_small_cin  : Cin <= 4  (conv_in with 1/3/4 image channels, first hint conv with 3 channels), Cout % 4 == 0.
//                     A thread owns one output pixel x a block of 32 output channels (accumulators in registers,
//                     weights [K][32] in smem read as warp-uniform float4 broadcasts).
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

// One thread = one output pixel x one block of COB output channels: the taps are loaded once and reused for all COB
// channels (weights are warp-uniform float4 broadcasts from smem), i.e. ~1.3 instructions per output value instead
// of the ~75 of a 4-channels-per-thread mapping, which made the first version of this kernel issue-bound at 160 us for
// conv_in (1 -> 32 @ 28x28, B = 1024) against a 20 us HBM time.
template <int COB>
__global__ void __launch_bounds__(128)
conv_small_cin_kernel(const __grid_constant__ SmallArgs a) {
  const cnb_conv_params& p = a.p;
  const int K = a.K;                       // ntaps * Cin
  const int n0 = blockIdx.y * COB;
  const float* in = reinterpret_cast<const float*>(p.in);
  extern __shared__ __align__(16) float wsm[];   // [K][COB]
  for (int i = threadIdx.x; i < K * COB; i += blockDim.x) {
    const int k = i / COB, o = i - k * COB;
    wsm[i] = __ldg(p.weight + (size_t)(n0 + o) * K + k);
  }
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const int OHW = p.OH * p.OW;
  const int b = m / OHW;
  const int rem = m - b * OHW;
  const int oy = rem / p.OW;
  const int ox = rem - oy * p.OW;
  const int iy0 = oy * p.stride, ix0 = ox * p.stride;
  const int Cin = p.Cin;
  float acc[COB];
#pragma unroll
  for (int j = 0; j < COB; ++j) acc[j] = 0.f;
  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int iy = iy0 + p.dy[tap], ix = ix0 + p.dx[tap];
    if ((unsigned)iy >= (unsigned)p.H || (unsigned)ix >= (unsigned)p.W) continue;
    const float* src = in + ((size_t)(b * p.H + iy) * p.W + ix) * p.ldi + p.in_coff;
    const float4* wt = reinterpret_cast<const float4*>(wsm + (size_t)tap * Cin * COB);
    for (int c = 0; c < Cin; ++c) {
      const float x = __ldg(src + c);
#pragma unroll
      for (int j = 0; j < COB / 4; ++j) {
        const float4 w = wt[c * (COB / 4) + j];
        acc[4 * j + 0] = fmaf(x, w.x, acc[4 * j + 0]); acc[4 * j + 1] = fmaf(x, w.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(x, w.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(x, w.w, acc[4 * j + 3]);
      }
    }
  }
  const size_t pix = ((size_t)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
#pragma unroll
  for (int j = 0; j < COB / 4; ++j) {
    const int n = n0 + 4 * j;
    float4 o = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    if (p.bias) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
    }
    if (p.temb) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.temb + (size_t)(p.temb_per_sample ? b : 0) * p.temb_ld + n));
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
  return p->Cin <= 4 && p->in_dtype == 0 && p->Cout % 4 == 0 && p->Cout >= 4 && K <= 64 &&
         (p->Cout % 32 == 0 || p->Cout == 16 || p->Cout == 8 || p->Cout == 4) && p->ldo % 4 == 0 && p->out_coff % 4 == 0 &&
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
    const int gx = ceil_div(a.M, 128);
    if (p->Cout % 32 == 0)
      conv_small_cin_kernel<32><<<dim3(gx, p->Cout / 32), 128, (size_t)a.K * 32 * sizeof(float), st>>>(a);
    else if (p->Cout == 16)
      conv_small_cin_kernel<16><<<dim3(gx, 1), 128, (size_t)a.K * 16 * sizeof(float), st>>>(a);
    else if (p->Cout == 8)
      conv_small_cin_kernel<8><<<dim3(gx, 1), 128, (size_t)a.K * 8 * sizeof(float), st>>>(a);
    else
      conv_small_cin_kernel<4><<<dim3(gx, 1), 128, (size_t)a.K * 4 * sizeof(float), st>>>(a);
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
