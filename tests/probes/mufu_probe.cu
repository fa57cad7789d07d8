// Does a stream of MUFU.EX2 from one warp starve the other warps of the same SM sub-partition?
// Block = 8 warps (2 per SMSP: warps w and w+4 share SMSP w%4).  Warps 0..3 run kind A, warps 4..7 run kind B.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// kind: 0 = idle, 1 = independent MUFU x N, 2 = independent FFMA x N, 3 = dependent FFMA chain, 4 = mixed FFMA+MUFU+pack like softmax
__device__ float run_kind(int kind, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f;
  if (kind == 1) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2f(a[i]);
    }
  } else if (kind == 2) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
    }
  } else if (kind == 3) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[0]) : "f"(0.999f), "f"(0.001f));
    }
  } else if (kind == 5) {          // ex2.approx.f16x2: two exponentials per instruction
    unsigned h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = __float_as_uint(a[i]) & 0x3BFF3BFFu;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(h[i]);
  } else if (kind == 6) {          // tanh.approx.f32
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  } else if (kind == 7) {          // cvt.rn.f16x2.f32 (SASS F2FP.F16.F32.PACK_AB): which pipe, what rate?
    unsigned h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        a[i] = __uint_as_float(h[i] | 0x3F000000u);
      }
    }
  } else if (kind == 8) {          // the softmax inner pattern: 2 x (FFMA, MUFU.EX2) + 1 x F2FP per score pair
    unsigned h[16];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-0.001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 1]) : "f"(0.999f), "f"(-0.001f));
        a[i] = ex2f(a[i]);
        a[i + 1] = ex2f(a[i + 1]);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[i + 1]));
      }
#pragma unroll
      for (int i = 0; i < 16; i += 2) a[i] += __uint_as_float(h[i] & 0x007F0000u);
    }
  } else if (kind == 9) {          // the same without the pack (2 x (FFMA, MUFU.EX2))
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-0.001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 1]) : "f"(0.999f), "f"(-0.001f));
        a[i] = ex2f(a[i]);
        a[i + 1] = ex2f(a[i + 1]);
      }
    }
  } else if (kind == 10) {         // mma.sync.m16n8k16 f16 -> f32 (SASS HMMA.16816.F32), four independent accumulators
    float c[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = a[j];
    const unsigned fa0 = __float_as_uint(a[4]), fa1 = __float_as_uint(a[5]), fa2 = __float_as_uint(a[6]), fa3 = __float_as_uint(a[7]);
    const unsigned fb0 = __float_as_uint(a[8]), fb1 = __float_as_uint(a[9]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[i & 3][0]), "+f"(c[i & 3][1]), "+f"(c[i & 3][2]), "+f"(c[i & 3][3])
                     : "r"(fa0), "r"(fa1), "r"(fa2), "r"(fa3), "r"(fb0), "r"(fb1));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = c[j][0] + c[j][1] + c[j][2] + c[j][3];
  } else if (kind == 4) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(-0.001f));
        a[i] = ex2f(a[i]);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  return s;
}
__global__ void probe(int kindA, int kindB, int iters, long long* out, float* sink) {
  const int warp = threadIdx.x >> 5;
  const int kind = warp < 4 ? kindA : kindB;
  __syncthreads();
  const long long t0 = clock64();
  float s = run_kind(kind, iters, threadIdx.x * 1e-3f);
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = s;
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 4096);
  const int iters = 256;   // x16 instructions
  const char* nm[11] = {"idle", "MUFU", "FFMA-indep", "FFMA-chain", "FFMA+MUFU", "EX2.F16x2", "TANH", "F2FP", "softmax8", "softmax8-nopack", "HMMA.16816"};
  // kinds 8 / 9 issue 8 MUFU per 16-slot iteration: their cyc/instr column is per 1/16 of an iteration (x2 = per MUFU)
  for (int ka : {0, 1, 5, 7, 8, 9, 10}) for (int kb : {0, 1, 5, 6, 7, 8, 9, 10, 2}) {
    if (ka > 1 && ka != 7 && ka != 10 && kb != 0 && kb != ka) continue;
    if (const char* only = getenv("PROBE_ONLY_HMMA")) { if (only[0] == '1' && ka != 10 && kb != 10) continue; }
    probe<<<1, 256>>>(ka, kb, iters, d, sink);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("A=%-10s B=%-10s  A: %6.2f cyc/instr   B: %6.2f cyc/instr\n", nm[ka], nm[kb], h[0] / (double)(iters * 16) / (ka == 4 ? 2 : 1),
           h[4] / (double)(iters * 16) / (kb == 4 ? 2 : 1));
  }
  return 0;
}
