// Flash attention on the 5th-generation tensor core with S, P AND O resident in tensor memory
// (fp16 q|k|v [B, L, 3E] -> fp16 out [B, L, E]; replaces the core of nn.MultiheadAttention, unet_base.py:107).
//
// Why this shape of kernel.  At the head dims of this model (4 .. 64) a score costs 4 d tensor FLOPs but one exponential,
// so the roofline is the MUFU pipe (16 ex2 / clk / SM), not the tensor pipe.  A kernel reaches it only if a softmax
// thread spends its issue slots on nothing but  FFMA (scale - max) -> ex2 -> pack  and synchronises rarely:
//   * S = Q K^T is a tcgen05.mma (M 128 x N BK x K DL) into TMEM; a softmax thread owns ONE query row (its TMEM lane) and
//     reads the whole BK-key row with tcgen05.ld: no ldmatrix, no HMMA, no quad shuffles, row maximum without shuffles;
//   * P is written back as packed fp16 pairs OVER the S columns it came from (tcgen05.st) and is the A operand of the
//     second MMA straight from TMEM (O += P [V | 1], "TS" form): no shared-memory round trip, no proxy fence;
//   * NBUF (3) S buffers per CTA: Q K^T runs up to three tiles ahead of the softmax, so a softmax warp goes from one tile
//     to the next through one (already completed) mbarrier wait and the MMA issue / commit latencies (measured: ~250
//     cycles per issue, 300 - 450 until a commit is visible, ~150 to wake a parked waiter) are off its critical path;
//   * the ones column appended to V makes the tensor core accumulate the softmax denominator (column D of O) in fp32;
//   * lazy rescale: the reference maximum of a row only moves when a chunk exceeds it by more than 2^8 (P <= 256 is
//     harmless in fp16, the sums are fp32); only then the warp waits for the previous PV MMA and scales its 32 rows of O
//     (and the part of the P row it already wrote) in TMEM.  After the first tile this is rare;
//   * POLY of every 32 exponentials are evaluated on the FMA pipe (Cody-Waite split + cubic) to go BELOW the MUFU roofline.
//
//   CTA = (sample b, head h, 128 query rows); two CTAs per SM (256 TMEM columns, ~40 KB of shared memory each); 8 warps:
//     warp 0      TMA producer: Q tile once, K tiles (3-stage ring), raw V tiles (2 stages)          (one thread)
//     warp 1      S issuer: S_j = Q K_j^T, NBUF tiles ahead                                             (one thread)
//     warp 3      O issuer: O += P_j [V_j | 1]; its commit hands the S buffer back to warp 1            (one thread)
//     warp 2      V transposer: [BK keys][DL] (TMA, no swizzle) -> V^T atoms [NV rows][64 keys], K-major SWIZZLE_128B;
//                 row D of V^T is all ones (written once)
//     warps 4-7   softmax, thread = query row
//   Head dims that are not a TMA / UMMA granule: DL >= D channels are loaded (D = 4 / 8: the aligned group of 16 that
//   contains the head; D = 24 / 48: 32 / 64 channels from the head's first) and the softmax threads zero the foreign
//   channels of their Q row in shared memory before the first MMA, so Q_masked K^T = Q_h K_h^T exactly.
//   TMEM (256 columns): S_0 .. S_{NBUF-1}, BK columns each | O, NV = round16(D + 1) columns.
//   Every mbarrier wait is bounded (tc_common.cuh) and latches the device error flag instead of hanging.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace cnb {
namespace atm {

using namespace tc;

constexpr int BQ = 128;            // query rows per CTA (UMMA M)
constexpr int NUM_THREADS = 256;
constexpr int KS = 3;              // K ring stages
constexpr float LAZY = 8.0f;       // log2 headroom before a row's reference maximum moves

struct alignas(64) Args {
  CUtensorMap map_q;               // [B*L, 3E] fp16, box {DL, 128}, swizzle = row bytes
  CUtensorMap map_k;               // box {DL, BK}, swizzle = row bytes
  CUtensorMap map_v;               // box {DL, BK}, no swizzle (read by the transposer warp)
  __half* out;
  int L, E;
  float scale_log2;
};

template <int DL, int D, int BK, int NBUF>
struct Cfg {
  static_assert(DL == 16 || DL == 32 || DL == 64, "loaded channels per row = one swizzle span");
  static_assert(D <= DL && BK % 16 == 0 && BK >= 48 && BK <= 128, "tile");
  static constexpr int RB = DL * 2;                          // bytes per Q / K / raw-V row
  static constexpr int NV = (D + 1 + 15) / 16 * 16;          // PV MMA N: D value columns, the ones column, zero pad
  static constexpr int NA = (BK + 31) / 32;                  // 32-key V^T atoms per tile (fp32 values: 32 keys = 128 bytes)
  static constexpr int Q_TILE = BQ * RB;
  static constexpr int K_TILE = BK * RB;
  static constexpr int VT_ATOM = NV * 128;
  static constexpr int VT_TILE = NA * VT_ATOM;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = (OFF_Q + Q_TILE + 1023) / 1024 * 1024;
  static constexpr int OFF_VR = (OFF_K + KS * K_TILE + 1023) / 1024 * 1024;
  static constexpr int OFF_VT = (OFF_VR + 2 * K_TILE + 1023) / 1024 * 1024;
  static constexpr int OFF_BAR = OFF_VT + 2 * VT_TILE;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;         // + alignment slack
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = NBUF * BK;
  static_assert(NBUF == 2 || NBUF == 3, "S buffers");
  static_assert(O_COL + NV <= TMEM_COLS, "TMEM budget: NBUF S buffers and O in 256 columns");
  static constexpr bool MASK_Q = D != DL;
};

// ---- PTX wrappers local to this kernel ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]   (A = P: fp32 bit patterns read as tf32, one column per key, lane = row; K = 8)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// 2^x on the FMA / ALU pipes: round-to-nearest split x = n + f, |f| <= 0.5, cubic minimax for 2^f (relative error
// 7.5e-5: below the fp16 rounding of P), exponent added to the bit pattern.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -60.f);
  const float t = x + 12582912.f;                           // 1.5 * 2^23: low mantissa bits = round(x)
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05517167f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One chunk of W (16 or 32) scores of this thread's row, in place: raw scores r[] -> P = 2^(s c - m c) as fp32 (the
// second MMA reads them as tf32).  `vk` = number of valid keys in the chunk (MASK only).
template <int W, bool MASK>
__device__ __forceinline__ void exp_inplace(uint32_t* r, float c, float mc, int vk) {
#pragma unroll
  for (int i = 0; i < W; ++i) {
    float e = ex2f(fmaf(__uint_as_float(r[i]), c, -mc));
    if (MASK && i >= vk) e = 0.f;
    r[i] = __float_as_uint(e);
  }
}

// ---- ordered instruction stream for the steady-state tiles ------------------------------------------------------------
// One warp issues in order; a MUFU.EX2 occupies its pipe for 8 cycles per warp instruction, so a softmax thread that
// runs "32 FFMA, 32 MUFU" back to back idles its issue slot 7 cycles out of 8 during the exponentials and then idles the
// MUFU pipe during everything else (measured: 45 % XU utilisation, ncu r2_03).  Here the exponentials of chunk c are
// interleaved, instruction by instruction, with the row-maximum and the scaling of chunk c+1 (whose TMEM load was
// issued before): `asm volatile` keeps the order through nvcc, and the chunk's own maximum test moves off the MUFU path.
__device__ __forceinline__ void v_ex2(float& x) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ void v_fma(float& x, float c, float nmc) {
  asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(c), "f"(nmc));
}
__device__ __forceinline__ void v_max3(float& t, float a, float b) {
  asm volatile("max.ftz.f32 %0, %0, %1, %2;" : "+f"(t) : "f"(a), "f"(b));
}
__device__ __forceinline__ uint32_t v_pack(float lo, float hi) {
  uint32_t d;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void v_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exp + pack of the current chunk (x = scaled scores, WC of them) interleaved with max + scale of the next chunk (raw
// scores r_next, WN of them, loaded by a tcgen05.ld that is still in flight on entry).  Returns the next chunk's maximum
// (raw units); r_next holds its scaled scores on exit.  POLY of every 8 exponentials run on the FMA pipe.
// filler pairs of the next chunk processed up to and including filler step k (1-based) out of `steps`: an even spread
__host__ __device__ constexpr int fill_cum(int k, int steps, int np) {
  return k <= 0 ? 0 : (k >= steps ? np : (np * k + steps - 1) / steps);
}

// Exponentials of the current chunk (x = scaled scores, WC of them, replaced by P in place) interleaved with max +
// scale of the next chunk (raw scores rn, WN of them, whose tcgen05.ld is still in flight on entry; scaled in place).
// P stays fp32: nothing consumes a MUFU result before the chunk's tcgen05.st, so the in-order warp never waits on the
// MUFU latency inside the chunk (with fp16 P every F2FP / PRMT sat two MUFU slots behind its operands and stalled on the
// short scoreboard, ncu r2_07).  POLY of every 4 pairs run on the FMA pipe.  Returns the next chunk's maximum (raw units).
template <int WC, int WN, int POLY>
__device__ __forceinline__ float exp_fill(float* x, float* rn, float c, float nmc) {
  constexpr int NS = WC / 2;                                 // steps = pairs of the current chunk
  constexpr int LAG = 2;                                     // pairs queued on the MUFU before waiting for the next chunk's load
  constexpr int NP = WN / 2;                                 // pairs of the next chunk
  constexpr int FS = NS - LAG;                               // filler steps (after the load wait)
  float t = -INFINITY;
#pragma unroll
  for (int sidx = 0; sidx < NS; ++sidx) {
    const int g = sidx & 3;                                  // position of the pair inside a group of 8 scores
    if ((POLY > 0 && g == 1) || (POLY > 1 && g == 3)) {
      x[2 * sidx] = ex2_poly(x[2 * sidx]);                   // this pair on the FMA pipe: no MUFU slot
      x[2 * sidx + 1] = ex2_poly(x[2 * sidx + 1]);
    } else {
      v_ex2(x[2 * sidx]);
      v_ex2(x[2 * sidx + 1]);
    }
    if (WN > 0) {
      if (sidx == LAG - 1) v_ld_wait();
      if (sidx >= LAG) {
        const int k = sidx - LAG + 1;
#pragma unroll
        for (int q = fill_cum(k - 1, FS, NP); q < fill_cum(k, FS, NP); ++q) {
          v_max3(t, rn[2 * q], rn[2 * q + 1]);
          v_fma(rn[2 * q], c, nmc);
          v_fma(rn[2 * q + 1], c, nmc);
        }
      }
    }
  }
  return t;
}

template <int W, bool MASK>
__device__ __forceinline__ float chunk_max(const uint32_t* r, int vk) {
  if (!MASK) {
    float t0 = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
    float t1 = fmax3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
#pragma unroll
    for (int i = 6; i + 3 < W; i += 4) {
      t0 = fmax3(t0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      t1 = fmax3(t1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
    // W = 16: indices 6..13 consumed by the loop (i = 6, 10), 14, 15 left; W = 32: i = 6 .. 26, then 30, 31 left
    return fmax3(t0, t1, fmaxf(__uint_as_float(r[W - 2]), __uint_as_float(r[W - 1])));
  }
  float t = -INFINITY;
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (i < vk) t = fmaxf(t, __uint_as_float(r[i]));
  return t;
}

// State of one softmax thread (query row) that the tile routines share.
struct SoftCtx {
  uint32_t t_o;                 // TMEM address of this warp's lanes, O columns
  uint64_t* pv_done;            // [2]
  volatile int* abort_flag;
  float c, lazy;                // scale * log2 e; headroom in raw score units
  float m;                      // reference maximum of the row (raw score units)
};

// Rare path: a chunk exceeded the row's reference maximum by more than the headroom.  Moves the reference of the lanes
// that need it, waits for the PV MMA of the previous tile, scales this warp's rows of O and the `nwritten` 16-column P
// granules of the current tile that were written against the old reference.  Warp-uniform control flow (tcgen05.ld / st
// are .sync.aligned); returns false if a bounded wait expired.
template <int NV>
__device__ __forceinline__ bool rescale_rows(SoftCtx& cx, float tm, int j, uint32_t t_s, int nwritten) {
  const float m_new = tm > cx.m + cx.lazy ? tm : cx.m;
  const float f = ex2f((cx.m - m_new) * cx.c);             // 1 for the lanes that keep their reference
  cx.m = m_new;
  if (j > 0) {
    if (!mbar_wait_parked(&cx.pv_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1), cx.abort_flag)) return false;
    tc_fence_after();
#pragma unroll
    for (int cb = 0; cb < NV; cb += 16) {
      uint32_t o[16];
      tmem_ld16_nowait(cx.t_o + cb, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
      tmem_st16(cx.t_o + cb, o);
    }
  }
  for (int pc = 0; pc < nwritten; ++pc) {                  // 16-column granules of fp32 P
    uint32_t o[16];
    tmem_ld16_nowait(t_s + (uint32_t)(pc * 16), o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
    tmem_st16(t_s + (uint32_t)(pc * 16), o);
  }
  tmem_st_wait();
  return true;
}

// One chunk step of a FULL tile (every key valid): exponentials of chunk CH interleaved with max + scale of chunk CH + 1.
template <int BK, int NV, int POLY, int CH>
__device__ __forceinline__ bool full_tile_step(SoftCtx& cx, int j, uint32_t t_s, float* cur, float* nxt, float& nmc) {
  constexpr int NCH = (BK + 31) / 32;
  if constexpr (CH < NCH) {
    constexpr int WC = (BK - 32 * CH) >= 32 ? 32 : 16;
    constexpr int WN = CH + 1 < NCH ? ((BK - 32 * (CH + 1)) >= 32 ? 32 : 16) : 0;
    if constexpr (WN == 32) tmem_ld32(t_s + (uint32_t)(32 * (CH + 1)), reinterpret_cast<uint32_t*>(nxt));
    if constexpr (WN == 16) tmem_ld16_nowait(t_s + (uint32_t)(32 * (CH + 1)), reinterpret_cast<uint32_t*>(nxt));
    const float tn = exp_fill<WC, WN, POLY>(cur, nxt, cx.c, nmc);
    if constexpr (WC == 32) tmem_st32(t_s + (uint32_t)(32 * CH), reinterpret_cast<uint32_t*>(cur));
    else tmem_st16(t_s + (uint32_t)(32 * CH), reinterpret_cast<uint32_t*>(cur));
    if constexpr (WN > 0) {
      if (__any_sync(0xffffffffu, tn > cx.m + cx.lazy)) {
        const float m_old = cx.m;
        if (!rescale_rows<NV>(cx, tn, j, t_s, 2 * (CH + 1))) return false;   // chunks 0 .. CH are 32 columns each
        const float dlt = (m_old - cx.m) * cx.c;             // the next chunk was scaled against the old reference
#pragma unroll
        for (int i = 0; i < WN; ++i) nxt[i] += dlt;
        nmc = -cx.m * cx.c;
      }
      return full_tile_step<BK, NV, POLY, CH + 1>(cx, j, t_s, nxt, cur, nmc);
    }
  }
  return true;
}

template <int BK, int NV, int POLY>
__device__ __forceinline__ bool full_tile(SoftCtx& cx, int j, uint32_t t_s) {
  float xa[32], xb[32];
  tmem_ld32(t_s, reinterpret_cast<uint32_t*>(xa));
  tmem_ld_wait();
  const float tm = chunk_max<32, false>(reinterpret_cast<uint32_t*>(xa), 32);
  if (j == 0) {
    cx.m = tm;
  } else if (__any_sync(0xffffffffu, tm > cx.m + cx.lazy)) {
    if (!rescale_rows<NV>(cx, tm, j, t_s, 0)) return false;
  }
  float nmc = -cx.m * cx.c;
#pragma unroll
  for (int i = 0; i < 32; ++i) xa[i] = fmaf(xa[i], cx.c, nmc);
  return full_tile_step<BK, NV, POLY, 0>(cx, j, t_s, xa, xb, nmc);
}

template <int DL, int D, int BK, int NBUF, int POLY>
__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tmem_kernel(const __grid_constant__ Args a) {
  using C = Cfg<DL, D, BK, NBUF>;
  constexpr int RB = C::RB, NV = C::NV;
  constexpr uint32_t IDESC_S = make_idesc(BQ, BK, true);
  constexpr uint32_t IDESC_PV = make_idesc(BQ, NV, false);   // kind::tf32: P (fp32 in TMEM) x V^T (fp32 in smem)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = raw_addr + pad;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_ready = bars + 1;             // Q rows masked by the softmax warps (MASK_Q only)
  uint64_t* k_full = bars + 2;              // [KS]
  uint64_t* k_empty = k_full + KS;          // [KS]
  uint64_t* v_full = k_empty + KS;          // [2] raw V landed (TMA)
  uint64_t* v_empty = v_full + 2;           // [2] raw V consumed by the transposer
  uint64_t* vt_full = v_empty + 2;          // [2] V^T written
  uint64_t* s_full = vt_full + 2;           // [NBUF] S buffer written by the MMA
  uint64_t* p_full = s_full + NBUF;         // [NBUF] P written over S by the active softmax warps
  uint64_t* s_free = p_full + NBUF;         // [NBUF] PV MMAs that read the P in this buffer finished (tcgen05.commit)
  uint64_t* pv_done = s_free + NBUF;        // [2] PV MMAs of the tile finished: V^T stage free, O readable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BQ;
  const int NT = (L + BK - 1) / BK;                        // key tiles
  const int row0 = b * L;                                   // first row of this sample in the [B*L, 3E] matrix
  const int nact = min(4, (L - q0 + 31) >> 5);              // softmax warps with at least one real query row
  // channel group of this head: xg = first loaded channel, c0 = offset of the head inside the loaded group
  const int xg = (DL % D == 0) ? (h * D / DL) * DL : h * D;
  const int c0 = h * D - xg;
  const int vlast = L - (NT - 1) * BK;                      // valid keys of the last tile (1 .. BK)

  if (tid == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_ready, (uint32_t)nact);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&vt_full[i], 1);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], (uint32_t)nact);
      mbar_init(&s_free[i], 1);
    }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  // rows D .. NV-1 of every V^T atom are constant: the ones row, then zeros (uniform rows: swizzle-free)
  if (warp == 2) {
    constexpr int ROWS = NV - D;
    for (int i = lane; i < 2 * C::NA * ROWS * 8; i += 32) {   // (atom of either stage, row, 16-byte chunk)
      const int chunk = i & 7, row = (i >> 3) % ROWS, atom = (i >> 3) / ROWS;
      const uint32_t v = row == 0 ? 0x3F800000u : 0u;          // 1.0f
      *reinterpret_cast<uint4*>(smem + C::OFF_VT + atom * C::VT_ATOM + (D + row) * 128 + chunk * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                               // every global access of this kernel comes after this point
  pdl_trigger();

  if (warp == 0) {
    // ======================= TMA producer (one thread) =======================
    if (lane == 0) {
      mbar_expect_tx(q_full, (uint32_t)C::Q_TILE);
      tma_load_2d(sbase + C::OFF_Q, &a.map_q, q_full, xg, row0 + q0);
      const int xk = a.E + xg, xv = 2 * a.E + xg;
      int st = 0;
      uint32_t ph = 1;                                       // parity of the k_empty phase that frees stage st (first lap: free)
      for (int j = 0; j < NT; ++j) {
        if (j >= KS && !mbar_wait_parked(&k_empty[st], ph, abort_flag)) break;
        mbar_expect_tx(&k_full[st], (uint32_t)C::K_TILE);
        tma_load_2d(sbase + C::OFF_K + st * C::K_TILE, &a.map_k, &k_full[st], xk, row0 + j * BK);
        if (++st == KS) { st = 0; ph ^= 1; }
        const int vs = j & 1;
        if (j >= 2 && !mbar_wait_parked(&v_empty[vs], (uint32_t)(((j >> 1) - 1) & 1), abort_flag)) break;
        mbar_expect_tx(&v_full[vs], (uint32_t)C::K_TILE);
        tma_load_2d(sbase + C::OFF_VR + vs * C::K_TILE, &a.map_v, &v_full[vs], xv, row0 + j * BK);
      }
    }
  } else if (warp == 1) {
    // ======================= S issuer (one thread): S_j = Q K_j^T, NBUF tiles ahead of the softmax =======================
    if (lane == 0 && mbar_wait_parked(C::MASK_Q ? q_ready : q_full, 0u, abort_flag)) {
      tc_fence_after();
      const uint64_t qdesc = make_desc_kmajor<RB>(sbase + C::OFF_Q);
      const uint64_t kdesc0 = make_desc_kmajor<RB>(sbase + C::OFF_K);
      int kst = 0, buf = 0;
      uint32_t kph = 0, fph = 1;                             // s_free parity of the next RE-use of `buf` (first use: free)
      for (int j = 0; j < NT; ++j) {
        // the buffer holds P(j - NBUF) until PV(j - NBUF) has read it: that MMA and this one differ in shape and
        // accumulator, the tensor pipe does not order them, so the hand-over goes through PV's commit (s_free)
        if (j >= NBUF && !mbar_wait_parked(&s_free[buf], fph, abort_flag)) break;
        if (!mbar_wait_parked(&k_full[kst], kph, abort_flag)) break;
        tc_fence_after();
        const uint64_t kdesc = kdesc0 + (uint64_t)((kst * C::K_TILE) >> 4);
#pragma unroll
        for (int k = 0; k < DL / 16; ++k)
          umma<true>(tmem_base + (uint32_t)(buf * BK), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), IDESC_S, k ? 1u : 0u);
        umma_commit(&k_empty[kst]);
        umma_commit(&s_full[buf]);
        if (++kst == KS) { kst = 0; kph ^= 1; }
        if (++buf == NBUF) { buf = 0; fph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == 3) {
    // ======================= O issuer (one thread): O += P_j [V_j | 1], P read from tensor memory =======================
    if (lane == 0) {
      const uint64_t vdesc0 = make_desc_kmajor<128>(sbase + C::OFF_VT);
      int buf = 0;
      uint32_t bph = 0;                                      // parity of the current use of S buffer `buf`
      for (int j = 0; j < NT; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        if (!mbar_wait_parked(&p_full[buf], bph, abort_flag)) break;
        if (!mbar_wait_parked(&vt_full[st], ph, abort_flag)) break;
        tc_fence_after();
        const int nkb = j < NT - 1 ? BK / 8 : (vlast + 7) >> 3;        // 8-key blocks (tf32 K) with at least one valid key
        const uint32_t p_col = tmem_base + (uint32_t)(buf * BK);
        const uint64_t vdesc_t = vdesc0 + (uint64_t)((st * C::VT_TILE) >> 4);
#pragma unroll
        for (int kb = 0; kb < BK / 8; ++kb) {
          if (kb < nkb)
            umma_ts(tmem_base + (uint32_t)C::O_COL, p_col + (uint32_t)(kb * 8),
                    vdesc_t + (uint64_t)(((kb >> 2) * C::VT_ATOM + (kb & 3) * 32) >> 4), IDESC_PV, (j | kb) ? 1u : 0u);
        }
        umma_commit(&s_free[buf]);
        umma_commit(&pv_done[st]);
        if (++buf == NBUF) { buf = 0; bph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == 2) {
    // ======================= V transposer: lane = one key of each 32-key atom, fp16 -> fp32 =======================
    const int chunk = lane >> 2, within = (lane & 3) * 4;   // 16-byte chunk of the V^T row (4 keys), byte offset inside it
    for (int j = 0; j < NT; ++j) {
      const int vs = j & 1, st = j & 1;
      if (!mbar_wait_parked(&v_full[vs], (uint32_t)((j >> 1) & 1), abort_flag)) break;
      if (j >= 2 && !mbar_wait_parked(&pv_done[st], (uint32_t)(((j >> 1) - 1) & 1), abort_flag)) break;
#pragma unroll
      for (int atom = 0; atom < C::NA; ++atom) {
        const int key = atom * 32 + lane;
        if (key < BK) {
          const uint8_t* src = smem + C::OFF_VR + vs * C::K_TILE + key * RB + c0 * 2;
          uint8_t* vt = smem + C::OFF_VT + st * C::VT_TILE + atom * C::VT_ATOM + within;
          if constexpr (D % 8 == 0) {
#pragma unroll
            for (int v = 0; v < D / 8; ++v) {
              const uint4 x0 = *reinterpret_cast<const uint4*>(src + v * 16);
              const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int dd = v * 8 + e;                      // V^T row
                const __half2 hp = *reinterpret_cast<const __half2*>(&w0[e >> 1]);
                const float val = (e & 1) ? __high2float(hp) : __low2float(hp);
                *reinterpret_cast<float*>(vt + dd * 128 + ((chunk ^ (dd & 7)) << 4)) = val;
              }
            }
          } else {                                             // D = 4: 8 bytes per key
            const uint2 x0 = *reinterpret_cast<const uint2*>(src);
            const uint32_t w0[2] = {x0.x, x0.y};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 hp = *reinterpret_cast<const __half2*>(&w0[e >> 1]);
              const float val = (e & 1) ? __high2float(hp) : __low2float(hp);
              *reinterpret_cast<float*>(vt + e * 128 + ((chunk ^ (e & 7)) << 4)) = val;
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&vt_full[st]);
        mbar_arrive(&v_empty[vs]);
      }
    }
  } else if (warp >= 4 && warp - 4 < nact) {
    // ======================= softmax: thread = query row =======================
    const int q = warp - 4;
    const int row = q * 32 + lane;                          // TMEM lane = row of the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_o = t_lane + (uint32_t)C::O_COL;
    const float c = a.scale_log2;
    const float lazy = LAZY / c;                            // headroom in raw score units
    bool dead = false;

    if constexpr (C::MASK_Q) {
      // zero the channels of this Q row that belong to other heads: Q_masked K^T = Q_h K_h^T
      if (mbar_wait_parked(q_full, 0u, abort_flag)) {
        constexpr int NC = RB / 16;                          // 16-byte chunks per row
        const int sw = (row >> (RB == 32 ? 2 : (RB == 64 ? 1 : 0))) & (NC - 1);
        uint8_t* qrow = smem + C::OFF_Q + row * RB;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          uint4* ptr = reinterpret_cast<uint4*>(qrow + ((j ^ sw) << 4));
          uint4 v = *ptr;
          uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = j * 8 + 2 * e;                    // channels ch, ch + 1 (c0 and D are even)
            if (ch < c0 || ch >= c0 + D) w[e] = 0u;
          }
          *ptr = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_ready);
      } else {
        dead = true;
      }
    }

    SoftCtx cx;
    cx.t_o = t_o;
    cx.pv_done = pv_done;
    cx.abort_flag = abort_flag;
    cx.c = c;
    cx.lazy = lazy;
    cx.m = 0.f;                                             // set by the first chunk
    float& m = cx.m;
    int buf = 0;
    uint32_t bph = 0;
    for (int j = 0; !dead && j < NT; ++j) {
      const uint32_t t_s = t_lane + (uint32_t)(buf * BK);
      if (!mbar_wait_parked(&s_full[buf], bph, abort_flag)) { dead = true; break; }
      tc_fence_after();
      const bool last = j == NT - 1;
      if (!last) {
        // steady state: straight-line, mask-free, exponentials interleaved with the next chunk's max / scale
        if (!full_tile<BK, NV, POLY>(cx, j, t_s)) { dead = true; break; }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
        if (++buf == NBUF) { buf = 0; bph ^= 1; }
        continue;
      }
      const int vk_tile = last ? vlast : BK;
      constexpr int NCH = (BK + 31) / 32;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        constexpr int dummy = 0;
        (void)dummy;
        const int col = ch * 32;
        const bool w16 = (BK - col) < 32;                    // tail chunk of 16 columns (BK = 112, 80)
        if (col >= vk_tile) break;                           // whole chunk beyond the sequence: the PV MMA skips it too
        uint32_t r[32];
        if (w16) {
          tmem_ld16_nowait(t_s + (uint32_t)col, r);
        } else {
          tmem_ld32(t_s + (uint32_t)col, r);
        }
        tmem_ld_wait();
        const int vk = vk_tile - col;                        // valid scores of this chunk (>= 1)
        const bool full = w16 ? vk >= 16 : vk >= 32;
        float tm;
        if (full) tm = w16 ? chunk_max<16, false>(r, 16) : chunk_max<32, false>(r, 32);
        else tm = w16 ? chunk_max<16, true>(r, vk) : chunk_max<32, true>(r, vk);
        if (j == 0 && ch == 0) {
          m = tm;
        } else if (__any_sync(0xffffffffu, tm > m + lazy)) {
          if (!rescale_rows<NV>(cx, tm, j, t_s, 2 * ch)) { dead = true; break; }
        }
        const float mc = m * c;
        if (w16) {
          if (full) exp_inplace<16, false>(r, c, mc, 16);
          else exp_inplace<16, true>(r, c, mc, vk);
          tmem_st16(t_s + (uint32_t)col, r);
        } else {
          if (full) exp_inplace<32, false>(r, c, mc, 32);
          else exp_inplace<32, true>(r, c, mc, vk);
          tmem_st32(t_s + (uint32_t)col, r);
        }
      }
      if (dead) break;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
      if (++buf == NBUF) { buf = 0; bph ^= 1; }
    }

    // ---- epilogue: out = O[:, :D] / O[:, D]
    if (!dead && mbar_wait_parked(&pv_done[(NT - 1) & 1], (uint32_t)(((NT - 1) >> 1) & 1), abort_flag)) {
      tc_fence_after();
      uint32_t o[NV];
#pragma unroll
      for (int cb = 0; cb < NV; cb += 16) tmem_ld16_nowait(t_o + cb, o + cb);
      tmem_ld_wait();
      if (q0 + row < L) {
        const float inv = 1.0f / __uint_as_float(o[D]);      // column D = sum of P = softmax denominator
        __half* op = a.out + ((size_t)(row0 + q0 + row) * a.E + h * D);
        if constexpr (D % 8 == 0) {
#pragma unroll
          for (int v = 0; v < D / 8; ++v) {
            uint4 w;
            w.x = pack2(__uint_as_float(o[8 * v + 0]) * inv, __uint_as_float(o[8 * v + 1]) * inv);
            w.y = pack2(__uint_as_float(o[8 * v + 2]) * inv, __uint_as_float(o[8 * v + 3]) * inv);
            w.z = pack2(__uint_as_float(o[8 * v + 4]) * inv, __uint_as_float(o[8 * v + 5]) * inv);
            w.w = pack2(__uint_as_float(o[8 * v + 6]) * inv, __uint_as_float(o[8 * v + 7]) * inv);
            *reinterpret_cast<uint4*>(op + 8 * v) = w;
          }
        } else {
          uint2 w;
          w.x = pack2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          w.y = pack2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          *reinterpret_cast<uint2*>(op) = w;
        }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

template <int DL, int D, int BK, int NBUF, int POLY>
static int launch(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  using C = Cfg<DL, D, BK, NBUF>;
  if (!g_encode) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    CNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) {
      set_error("attention_tmem: cuTensorMapEncodeTiled unavailable");
      return CNB_ERR_UNSUPPORTED;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(f);
  }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    CNB_CUDA(cudaFuncSetAttribute(attention_tmem_kernel<DL, D, BK, NBUF, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
  }
  Args a;
  memset(&a, 0, sizeof(a));
  const cuuint64_t dims[2] = {(cuuint64_t)3 * E, (cuuint64_t)B * L};
  const cuuint64_t strides[1] = {(cuuint64_t)3 * E * 2};
  const cuuint32_t box_q[2] = {(cuuint32_t)DL, (cuuint32_t)BQ};
  const cuuint32_t box_k[2] = {(cuuint32_t)DL, (cuuint32_t)BK};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz = DL == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (DL == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = g_encode(&a.map_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box_q, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = g_encode(&a.map_k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box_k, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = g_encode(&a.map_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box_k, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tmem: cuTensorMapEncodeTiled failed (%d): B=%d L=%d E=%d", (int)r, B, L, E);
    return CNB_ERR_CUDA;
  }
  a.out = reinterpret_cast<__half*>(out);
  a.L = L;
  a.E = E;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  dim3 grid(ceil_div(L, BQ), heads, B);
  CNB_CUDA(launch_pdl((long long)B * L * E, attention_tmem_kernel<DL, D, BK, NBUF, POLY>, grid, dim3(NUM_THREADS),
                      (size_t)C::TOTAL, st, a));
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

// Tile plan: NBUF S buffers of BK columns and O (NV columns) share the CTA's 256 TMEM columns.
//   3 buffers of 64 (80 when NV = 16) keys: Q K^T runs three tiles ahead of the softmax, the MMA issue chain is never on
//   the softmax warps' critical path;  2 buffers of up to 112 keys: fewer barrier rounds per row.
// CNB_ATTN_TMEM_NBUF / CNB_ATTN_TMEM_BK override the measured defaults.
template <int DL, int D, int POLY>
static int launch_bk(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  constexpr int NV = (D + 1 + 15) / 16 * 16;
  // two S buffers: 112 (NV <= 32), 96 (NV 48); head dims loaded as 64 channels (48, 64) stop at 64 keys so that two
  // CTAs (K ring + raw V + fp32 V^T stages) still fit one SM's shared memory
  constexpr int BK2 = DL == 64 ? 64 : ((256 - NV) / 2) / 16 * 16;
  constexpr int BK3 = DL == 64 ? 48 : ((256 - NV) / 3) / 16 * 16;   // 80 (NV 16), 64 (NV 32 / 48)
  static int bk_env = -1, nbuf_env = -1;
  if (bk_env < 0) {
    const char* e = getenv("CNB_ATTN_TMEM_BK");
    bk_env = e ? atoi(e) : 0;
    const char* n = getenv("CNB_ATTN_TMEM_NBUF");
    nbuf_env = n ? atoi(n) : 0;
  }
  int nbuf = nbuf_env ? nbuf_env : 2;
  if (L <= BK2) nbuf = 2;                                   // one tile: nothing to run ahead
  if (nbuf == 3) {
    int bk = bk_env && bk_env <= BK3 ? bk_env : BK3;
    if constexpr (BK3 >= 80) if (bk == 80) return launch<DL, D, 80, 3, POLY>(qkv, out, B, L, E, heads, st);
    if constexpr (BK3 >= 64) if (bk >= 64) return launch<DL, D, 64, 3, POLY>(qkv, out, B, L, E, heads, st);
    return launch<DL, D, 48, 3, POLY>(qkv, out, B, L, E, heads, st);
  }
  auto padded = [&](int bk) { return ceil_div(L, bk) * bk; };
  int bk = BK2;
  if (bk_env) bk = bk_env < BK2 ? bk_env : BK2;
  else {
    const int cands[3] = {96, 80, 64};                      // a smaller tile only if it saves more than 6 % of the padded keys
    for (int i = 0; i < 3; ++i)
      if (cands[i] < bk && padded(cands[i]) * 100 < padded(bk) * 94) bk = cands[i];
  }
  if constexpr (BK2 >= 112) if (bk == 112) return launch<DL, D, 112, 2, POLY>(qkv, out, B, L, E, heads, st);
  if constexpr (BK2 >= 96) if (bk == 96) return launch<DL, D, 96, 2, POLY>(qkv, out, B, L, E, heads, st);
  if (bk == 80) return launch<DL, D, 80, 2, POLY>(qkv, out, B, L, E, heads, st);
  return launch<DL, D, 64, 2, POLY>(qkv, out, B, L, E, heads, st);
}

template <int POLY>
static int launch_d(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  switch (E / heads) {
    case 4: return launch_bk<16, 4, POLY>(qkv, out, B, L, E, heads, st);
    case 8: return launch_bk<16, 8, POLY>(qkv, out, B, L, E, heads, st);
    case 16: return launch_bk<16, 16, POLY>(qkv, out, B, L, E, heads, st);
    case 24: return launch_bk<32, 24, POLY>(qkv, out, B, L, E, heads, st);
    case 32: return launch_bk<32, 32, POLY>(qkv, out, B, L, E, heads, st);
    case 48: return launch_bk<64, 48, POLY>(qkv, out, B, L, E, heads, st);
    case 64: return launch_bk<64, 64, POLY>(qkv, out, B, L, E, heads, st);
    default:
      set_error("attention_tmem: head dim %d not instantiated", E / heads);
      return CNB_ERR_UNSUPPORTED;
  }
}

}  // namespace atm

int attention_tmem_error_flag() { return tc_read_clear_error(); }

// default routing of cnb_attention_f16: CNB_ATTN_TMEM=0 falls back to the mma.sync kernel everywhere, =1 (default)
// uses this kernel for every supported shape with L >= CNB_ATTN_TMEM_MINL (default 96)
bool attention_tmem_default(int L) {
  static int on = -1, minl = 96;
  if (on < 0) {
    const char* e = getenv("CNB_ATTN_TMEM");
    on = e ? atoi(e) : 1;
    const char* m = getenv("CNB_ATTN_TMEM_MINL");
    if (m) minl = atoi(m);
  }
  return on != 0 && L >= minl;
}

bool attention_tmem_supported(const void* qkv, const void* out, int B, int L, int E, int heads) {
  if (heads <= 0 || E % heads) return false;
  const int d = E / heads;
  if (d != 4 && d != 8 && d != 16 && d != 24 && d != 32 && d != 48 && d != 64) return false;
  if (L < 1 || B > 65535 || heads > 65535) return false;
  if ((3 * E * 2) % 16 || (((uintptr_t)qkv | (uintptr_t)out) & 15)) return false;
  return true;
}

int attention_tmem(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  static int poly = -1;                                    // eighths of the exponentials evaluated on the FMA pipe
  if (poly < 0) {
    const char* e = getenv("CNB_ATTN_POLY");
    poly = e ? atoi(e) : 0;
  }
  switch (poly) {                                          // quarters of the exponential PAIRS evaluated on the FMA pipe
    case 1: return atm::launch_d<1>(qkv, out, B, L, E, heads, st);
    case 2: return atm::launch_d<2>(qkv, out, B, L, E, heads, st);
    default: return atm::launch_d<0>(qkv, out, B, L, E, heads, st);
  }
}

}  // namespace cnb
