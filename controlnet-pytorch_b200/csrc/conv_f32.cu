// fp32 implicit-GEMM convolution on CUDA cores (CNB_MODE_F32): the exact-arithmetic path ("fp32 mode", 1e-4
// gate of BASELINE.json) and the fallback for shapes the tcgen05 kernel does not take (Cin % 4 != 0: conv_in
// with 1/3 input channels; Cout < 16: conv_out).  GEMM view:  M = B*OH*OW pixels, N = Cout, K = ntaps*Cin,
// A gathered on the fly from the channels-last activation tensor (zero outside the image), W is [N][K].
//
// Tile BM x BN x 16, 256 threads, TM x TN register micro-tiles, register-prefetched global loads.
#include "common.cuh"

namespace cnb {

struct ConvArgs {
  cnb_conv_params p;
  int M, K;
};

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_igemm_f32_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_ELEMS = BM * BK / NT;   // scalars per thread per tile
  constexpr int B_ELEMS = (BN * BK + NT - 1) / NT;
  static_assert(A_ELEMS % 4 == 0, "A tile must split into float4 per thread");
  constexpr int A_VEC = A_ELEMS / 4;
  constexpr int B_VEC = (BN * BK / 4 + NT - 1) / NT;

  const cnb_conv_params& p = a.p;
  const float* p_in = reinterpret_cast<const float*>(p.in);
  const int M = a.M, K = a.K;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int OHW = p.OH * p.OW;
  const bool vecA = (p.Cin % 4 == 0) && (p.ldi % 4 == 0) && (p.in_coff % 4 == 0);
  const bool vecB = (K % 4 == 0);

  // ---- per-thread row descriptors for the vector A path (row is fixed across k-steps)
  int a_row[A_VEC], a_k4[A_VEC], a_iy0[A_VEC], a_ix0[A_VEC];
  long long a_base[A_VEC];
  bool a_ok[A_VEC];
#pragma unroll
  for (int i = 0; i < A_VEC; ++i) {
    int u = tid + i * NT;
    a_row[i] = u / (BK / 4);
    a_k4[i] = u % (BK / 4);
    int m = m0 + a_row[i];
    a_ok[i] = m < M;
    int mm = a_ok[i] ? m : 0;
    int b = mm / OHW;
    int r = mm - b * OHW;
    int oy = r / p.OW;
    int ox = r - oy * p.OW;
    a_iy0[i] = oy * p.stride;
    a_ix0[i] = ox * p.stride;
    a_base[i] = (long long)b * p.H * p.W;
  }

  float4 ra[A_VEC];
  float rb[B_VEC * 4];

  auto load_tile = [&](int k0) {
    if (vecA) {
#pragma unroll
      for (int i = 0; i < A_VEC; ++i) {
        int k = k0 + a_k4[i] * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_ok[i] && k < K) {
          int tap = k / p.Cin;
          int c = k - tap * p.Cin;
          int iy = a_iy0[i] + p.dy[tap];
          int ix = a_ix0[i] + p.dx[tap];
          if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) {
            const float* src = p_in + ((a_base[i] + (long long)iy * p.W + ix) * p.ldi + p.in_coff + c);
            v = __ldg(reinterpret_cast<const float4*>(src));
          }
        }
        ra[i] = v;
      }
    } else {
      float* raf = reinterpret_cast<float*>(ra);
#pragma unroll
      for (int i = 0; i < A_ELEMS; ++i) {
        int e = tid + i * NT;
        int row = e / BK, kk = e % BK;
        int m = m0 + row, k = k0 + kk;
        float v = 0.f;
        if (m < M && k < K) {
          int b = m / OHW;
          int r = m - b * OHW;
          int oy = r / p.OW, ox = r - (r / p.OW) * p.OW;
          int tap = k / p.Cin;
          int c = k - tap * p.Cin;
          int iy = oy * p.stride + p.dy[tap];
          int ix = ox * p.stride + p.dx[tap];
          if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
            v = __ldg(p_in + (((long long)b * p.H * p.W + (long long)iy * p.W + ix) * p.ldi + p.in_coff + c));
        }
        raf[i] = v;
      }
    }
    if (vecB) {
#pragma unroll
      for (int i = 0; i < B_VEC; ++i) {
        int u = tid + i * NT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < BN * BK / 4) {
          int n = n0 + u / (BK / 4);
          int k = k0 + (u % (BK / 4)) * 4;
          if (n < p.Cout && k < K) v = __ldg(reinterpret_cast<const float4*>(p.weight + (long long)n * K + k));
        }
        rb[i * 4 + 0] = v.x; rb[i * 4 + 1] = v.y; rb[i * 4 + 2] = v.z; rb[i * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_ELEMS; ++i) {
        int e = tid + i * NT;
        float v = 0.f;
        if (e < BN * BK) {
          int n = n0 + e / BK;
          int k = k0 + e % BK;
          if (n < p.Cout && k < K) v = __ldg(p.weight + (long long)n * K + k);
        }
        if (i < B_VEC * 4) rb[i] = v;
      }
    }
  };

  auto store_tile = [&]() {
    if (vecA) {
#pragma unroll
      for (int i = 0; i < A_VEC; ++i) {
        int kk = a_k4[i] * 4;
        As[kk + 0][a_row[i]] = ra[i].x;
        As[kk + 1][a_row[i]] = ra[i].y;
        As[kk + 2][a_row[i]] = ra[i].z;
        As[kk + 3][a_row[i]] = ra[i].w;
      }
    } else {
      const float* raf = reinterpret_cast<const float*>(ra);
#pragma unroll
      for (int i = 0; i < A_ELEMS; ++i) {
        int e = tid + i * NT;
        As[e % BK][e / BK] = raf[i];
      }
    }
    if (vecB) {
#pragma unroll
      for (int i = 0; i < B_VEC; ++i) {
        int u = tid + i * NT;
        if (u < BN * BK / 4) {
          int n = u / (BK / 4), kk = (u % (BK / 4)) * 4;
          Bs[kk + 0][n] = rb[i * 4 + 0];
          Bs[kk + 1][n] = rb[i * 4 + 1];
          Bs[kk + 2][n] = rb[i * 4 + 2];
          Bs[kk + 3][n] = rb[i * 4 + 3];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_ELEMS; ++i) {
        int e = tid + i * NT;
        if (e < BN * BK && i < B_VEC * 4) Bs[e % BK][e / BK] = rb[i];
      }
    }
  };

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nkb = (K + BK - 1) / BK;
  load_tile(0);
  for (int kb = 0; kb < nkb; ++kb) {
    store_tile();
    __syncthreads();
    if (kb + 1 < nkb) load_tile((kb + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- fused epilogue: + bias + time-embedding row + residual, optional SiLU
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    int b = m / OHW;
    int r = m - b * OHW;
    int oy = r / p.OW, ox = r - (r / p.OW) * p.OW;
    long long pix = ((long long)b * p.OHf + (oy * p.oy_mul + p.oy_add)) * p.OWf + (ox * p.ox_mul + p.ox_add);
    float* dst = reinterpret_cast<float*>(p.out) + pix * p.ldo + p.out_coff;
    const float* res = p.residual ? reinterpret_cast<const float*>(p.residual) + pix * p.ldr + p.res_coff : nullptr;
    const float* te = p.temb ? p.temb + (long long)(p.temb_per_sample ? b : 0) * p.temb_ld : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (te) v += __ldg(te + n);
      if (res) v += __ldg(res + n);
      if (p.act == 1) v = silu_f(v);
      dst[n] = v;
    }
  }
}

template <int BM, int BN, int TM, int TN>
static int launch(const ConvArgs& a, cudaStream_t st) {
  dim3 grid(ceil_div(a.M, BM), ceil_div(a.p.Cout, BN));
  conv_igemm_f32_kernel<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

int conv2d_f32(const cnb_conv_params* p, cudaStream_t st) {
  ConvArgs a;
  a.p = *p;
  a.M = p->B * p->OH * p->OW;
  a.K = p->ntaps * p->Cin;
  if (p->Cout > 32) return launch<128, 64, 8, 4>(a, st);
  if (p->Cout > 16) return launch<128, 32, 8, 2>(a, st);
  return launch<128, 16, 8, 1>(a, st);
}

}  // namespace cnb
