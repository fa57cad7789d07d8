// tcgen05 / TMEM flash attention for head dims 16 and 32 on long sequences (fp16 q|k|v [B, L, 3E] -> fp16 out [B, L, E]).
//
// Why a second attention kernel: the register-resident mma.sync kernel (attention_f16.cu) sits at the joint limit of the
// MUFU pipe (one ex2 per score) and the issue slots (ldmatrix, HMMA, quad shuffles, packing: ~6.7 instructions per
// score).  Here the contractions run on the 5th-gen tensor core with the accumulators in tensor memory and the softmax
// threads own half rows, so a score costs FFMA + ex2 + half a pack + 1/8 of a 16-byte smem store, and a configurable
// fraction of the exponentials moves to the FMA pipe (Cody-Waite + cubic) to go below the MUFU roofline.
//
//   CTA            = (batch b, head h, 128 query rows), 128-key tiles, ONE pass over the keys; 12 warps:
//     warp 0       TMA producer: Q tile, K tiles (ring), V tiles                                  (one thread)
//     warp 1       S = Q K^T issuer (tcgen05.mma M 128, N 128, K = D), one S buffer in TMEM       (one thread)
//     warp 2       O_hf += P_hf [V_hf | 1] issuer: two accumulators, one per 64-key half of a tile (one thread)
//     warp 3       V transposer: [128 keys][D] (TMA, no swizzle) -> two V^T atoms [D + 16][64 keys] in the K-major
//                  SWIZZLE_128B layout the PV MMA's B operand needs; row D of V^T is all ones, so column D of O is
//                  the softmax denominator accumulated by the tensor core in fp32
//     warps 4..11  softmax: warp w owns TMEM lane quarter w & 3 (32 query rows) and key half hf = (w - 4) >> 2 of
//                  every tile, read as two tcgen05.ld.32x32b.x32 batches; the S buffer is handed back after the
//                  second load, before its exponentials (the next Q K^T overlaps them).  Each (row, half) thread is an
//                  independent online softmax over its 64-key half tiles with its own reference maximum and its own
//                  accumulator O_hf (split-K), merged at the end.
//   tile size      measured (tests/probes/sync_probe.cu, tests/probes/atc5_trace.cu): a successful mbarrier try_wait costs ~90
//                  cycles, fences / syncwarp / arrive 25-40 each, so a softmax warp pays ~1000 cycles of
//                  synchronisation latency per barrier round against 8 MUFU cycles per score-instruction; 64 scores
//                  per thread and round keep the MUFU pipe fed with the four softmax warps an SM sub-partition holds.
//   lazy rescale   the reference maximum only moves when a batch maximum exceeds it by more than 2^8 (P <= 256 is
//                  harmless in fp16, sums are fp32); then the warp waits for the previous PV MMA, scales its 32 rows
//                  of O_hf in TMEM (tcgen05.ld / st), rescales the part of the P row it already wrote, and goes on.
//                  After the first tile this is rare: no correction step on the critical path, no second pass over K.
//   memory         Q, K: rows of D fp16 = one 32/64-byte swizzle row each (TMA 2-D tiles of the [B*L, 3E] matrix);
//                  P: two atoms of 128 x 64 fp16 per tile, written by the softmax threads in the K-major SWIZZLE_128B
//                  layout (PS stages); TMEM: 256 columns = 128 S columns + 2 x (D + 16) O columns; two CTAs per SM
//   sync           mbarriers only (full/empty pairs per ring, tcgen05.commit for MMA completion); every wait is bounded
//   measured       tests/probes/lat_probe.cu: one tcgen05.mma + commit costs the issuing thread ~250 cycles and a TMA issue
//                  ~300, hence one issuing thread per engine.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace cnb {
namespace atc5 {

using namespace tc;

constexpr int BQ = 128;          // query rows per CTA (UMMA M)
constexpr int BKEY = 128;        // keys per tile (UMMA N of S)
constexpr int NUM_THREADS = 384;
constexpr int NSM = 256;         // softmax threads
constexpr int KS = 3;            // K ring stages
constexpr int PS = 2;            // P / V^T stages
constexpr float LAZY = 8.0f;     // log2 headroom before the reference maximum moves

struct alignas(64) Args {
  CUtensorMap map_q;             // [B*L, 3E] fp16, box {D, 128 rows}, swizzle = D*2 bytes (Q and K tiles)
  CUtensorMap map_v;             // box {D, 128 rows}, no swizzle (read by the transposer warp)
  __half* out;
  int L, E, q_tiles;
  float scale_log2;
};

template <int D>
struct Smem {
  static constexpr int RBQ = D * 2;                        // bytes per Q / K row = swizzle span
  static constexpr int Q_TILE = BQ * RBQ;                  // 4 / 8 KB
  static constexpr int K_TILE = BKEY * RBQ;                // 4 / 8 KB
  static constexpr int NV = D + 16;                        // PV MMA N: D value columns + the ones column (+ zero pad)
  static constexpr int VT_ATOM = NV * 128;                 // V^T of 64 keys: NV rows x 128 bytes
  static constexpr int VT_TILE = 2 * VT_ATOM;
  static constexpr int P_ATOM = BQ * 128;                  // P of 64 keys: 128 rows x 128 bytes
  static constexpr int P_TILE = 2 * P_ATOM;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_TILE;             // KS stages
  static constexpr int OFF_VR = OFF_K + KS * K_TILE;       // raw V, 2 stages
  static constexpr int OFF_VT = (OFF_VR + 2 * K_TILE + 1023) / 1024 * 1024;   // V^T, PS stages
  static constexpr int OFF_P = (OFF_VT + PS * VT_TILE + 1023) / 1024 * 1024;  // P, PS stages
  static constexpr int OFF_MX = OFF_P + PS * P_TILE;       // reference maxima of key half 1: 128 floats
  static constexpr int OFF_BAR = OFF_MX + 512;
  static constexpr int TOTAL = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = BKEY;                       // O_0 at O_COL, O_1 at O_COL + NV
  static_assert(O_COL + 2 * NV <= TMEM_COLS, "TMEM budget");
};

#ifdef ATC5_TRACE
// development timeline (tests/probes/atc5_trace.cu): (tag, clock) pairs per role of one CTA in the middle of the grid
__device__ long long g_trace[8][512];
#define TR_DECL(role) const bool tr_on = blockIdx.x == 3 && blockIdx.y == 2 && blockIdx.z == 500; int tr_n = 0; const int tr_role = role;
#define TR(tag) do { if (tr_on && tr_n < 510) { g_trace[tr_role][tr_n++] = (tag); g_trace[tr_role][tr_n++] = clock64(); } } while (0)
#else
#define TR_DECL(role)
#define TR(tag)
#endif

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  tmem_ld16_nowait(taddr, r);
  tmem_ld16_nowait(taddr + 16, r + 16);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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

// P = exp2(S c - m c) for the 32 scores of this thread's key half -> four 16-byte chunks of the swizzled P row.
// POLY of every 8 exponentials run on the FMA pipe.  vh = valid scores of this half (MASK only).
template <bool MASK, int POLY>
__device__ __forceinline__ void exp_store(const uint32_t* r, float c, float mc, int vh, uint8_t* prow, int ch0, int rsw) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k0 = ch * 8 + 2 * i;
      const float x0 = fmaf(__uint_as_float(r[k0]), c, -mc), x1 = fmaf(__uint_as_float(r[k0 + 1]), c, -mc);
      float e0 = (POLY > 0 && (i == 1)) || (POLY > 2 && i == 3) ? ex2_poly(x0) : ex2f(x0);
      float e1 = (POLY > 1 && (i == 2)) || (POLY > 3 && i == 0) ? ex2_poly(x1) : ex2f(x1);
      if (MASK) {
        if (k0 >= vh) e0 = 0.f;
        if (k0 + 1 >= vh) e1 = 0.f;
      }
      pk[i] = pack2(e0, e1);
    }
    *reinterpret_cast<uint4*>(prow + (((ch0 + ch) ^ rsw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Multiply the 32 fp16 values of four 16-byte chunks of a P row by f (lazy-rescale fix-up of an already written batch).
__device__ __forceinline__ void rescale_p(uint8_t* prow, int ch0, int rsw, float f) {
  const __half2 f2 = __float2half2_rn(f);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4* ptr = reinterpret_cast<uint4*>(prow + (((ch0 + ch) ^ rsw) << 4));
    uint4 v = *ptr;
    __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) hv[i] = __hmul2(hv[i], f2);
    *ptr = v;
  }
}

template <int D, int POLY>
__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tc05_kernel(const __grid_constant__ Args a) {
  using S = Smem<D>;
  constexpr int RBQ = S::RBQ;
  constexpr uint32_t IDESC_S = make_idesc(BQ, BKEY, true);
  constexpr uint32_t IDESC_PV = make_idesc(BQ, S::NV, true);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sbase = raw_addr + pad;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;              // [KS]
  uint64_t* k_empty = k_full + KS;          // [KS]
  uint64_t* v_full = k_empty + KS;          // [2]  raw V landed (TMA)
  uint64_t* v_empty = v_full + 2;           // [2]  raw V consumed by the transposer
  uint64_t* vt_full = v_empty + 2;          // [PS] V^T written
  uint64_t* pv_done = vt_full + PS;         // [PS] PV MMAs of the tile finished: P and V^T stage free, O readable
  uint64_t* s_full = pv_done + PS;          // S buffer written by the MMA
  uint64_t* s_empty = s_full + 1;           // S buffer read out by the 8 softmax warps
  uint64_t* p_full = s_empty + 1;           // [PS] P written by the 8 softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + PS);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* smx = reinterpret_cast<float*>(smem + S::OFF_MX);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BQ;
  const int NT = (L + BKEY - 1) / BKEY;                   // key tiles
  const int row0 = b * L;                                  // first row of this sample in the [B*L, 3E] matrix

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, NSM / 32);
    for (int i = 0; i < PS; ++i) {
      mbar_init(&vt_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&p_full[i], NSM / 32);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  // the ones / zero rows of all V^T atoms (rows D .. D+15) are written once
  if (warp == 3) {
    for (int i = lane; i < PS * 2 * 16 * 8; i += 32) {      // (atom, row, 16-byte chunk)
      const int chunk = i & 7, row = (i >> 3) & 15, atom = i >> 7;
      const uint32_t v = row == 0 ? 0x3C003C00u : 0u;       // half2(1, 1) in row D, zeros below (uniform rows: swizzle-free)
      *reinterpret_cast<uint4*>(smem + S::OFF_VT + atom * S::VT_ATOM + (D + row) * 128 + chunk * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer (one thread) =======================
    if (lane == 0) {
      mbar_expect_tx(q_full, (uint32_t)S::Q_TILE);
      tma_load_2d(sbase + S::OFF_Q, &a.map_q, q_full, h * D, row0 + q0);
      const int xk = a.E + h * D, xv = 2 * a.E + h * D;
      TR_DECL(4)
      TR(0);
      int st = 0;
      uint32_t ph = 1;                                       // parity of the k_empty phase that frees stage st (first lap: free)
      for (int j = 0; j < NT; ++j) {
        if (j >= KS && !mbar_wait_parked(&k_empty[st], ph, abort_flag)) break;
        mbar_expect_tx(&k_full[st], (uint32_t)S::K_TILE);
        tma_load_2d(sbase + S::OFF_K + st * S::K_TILE, &a.map_q, &k_full[st], xk, row0 + j * BKEY);
        if (++st == KS) { st = 0; ph ^= 1; }
        const int vs = j & 1;
        if (j >= 2 && !mbar_wait_parked(&v_empty[vs], (uint32_t)(((j >> 1) - 1) & 1), abort_flag)) break;
        mbar_expect_tx(&v_full[vs], (uint32_t)S::K_TILE);
        tma_load_2d(sbase + S::OFF_VR + vs * S::K_TILE, &a.map_v, &v_full[vs], xv, row0 + j * BKEY);
        TR(1000 + j);
      }
    }
  } else if (warp == 1) {
    // ======================= S = Q K^T issuer (one thread) =======================
    if (lane == 0 && mbar_wait_parked(q_full, 0u, abort_flag)) {
      tc_fence_after();
      const uint64_t qdesc = make_desc_kmajor<RBQ>(sbase + S::OFF_Q);
      const uint64_t kdesc0 = make_desc_kmajor<RBQ>(sbase + S::OFF_K);
      int kst = 0;
      uint32_t kph = 0;
      TR_DECL(0)
      for (int j = 0; j < NT; ++j) {
        TR(1);
        if (j >= 1 && !mbar_wait_parked(s_empty, (uint32_t)((j - 1) & 1), abort_flag)) break;
        if (!mbar_wait_parked(&k_full[kst], kph, abort_flag)) break;
        tc_fence_after();
        TR(2);
        const uint64_t kdesc = kdesc0 + (uint64_t)((kst * S::K_TILE) >> 4);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma<true>(tmem_base, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), IDESC_S, k ? 1u : 0u);
        umma_commit(&k_empty[kst]);
        umma_commit(s_full);
        TR(3);
        if (++kst == KS) { kst = 0; kph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == 2) {
    // ======================= O_hf += P_hf [V_hf | 1] issuer (one thread) =======================
    if (lane == 0) {
      const uint64_t pdesc0 = make_desc_kmajor<128>(sbase + S::OFF_P);
      const uint64_t vdesc0 = make_desc_kmajor<128>(sbase + S::OFF_VT);
      int st = 0;
      uint32_t ph = 0;
      TR_DECL(1)
      for (int j = 0; j < NT; ++j) {
        TR(1);
        if (!mbar_wait_parked(&p_full[st], ph, abort_flag)) break;
        TR(2);
        if (!mbar_wait_parked(&vt_full[st], ph, abort_flag)) break;
        tc_fence_after();
        TR(3);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint64_t pdesc = pdesc0 + (uint64_t)((st * S::P_TILE + hf * S::P_ATOM) >> 4);
          const uint64_t vdesc = vdesc0 + (uint64_t)((st * S::VT_TILE + hf * S::VT_ATOM) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<true>(tmem_base + (uint32_t)(S::O_COL + hf * S::NV), pdesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k),
                       IDESC_PV, (j | k) ? 1u : 0u);
        }
        umma_commit(&pv_done[st]);
        TR(4);
        if (++st == PS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == 3) {
    // ======================= V transposer: lane = key pair of each 64-key atom =======================
    const int chunk = lane >> 2, within = (lane & 3) * 4;   // 16-byte chunk of the V^T row, byte offset inside it
    int st = 0;
    uint32_t ph = 1;
    TR_DECL(5)
    for (int j = 0; j < NT; ++j) {
      const int vs = j & 1;
      if (lane == 0) TR(1);
      if (!mbar_wait_parked(&v_full[vs], (uint32_t)((j >> 1) & 1), abort_flag)) break;
      if (j >= PS && !mbar_wait_parked(&pv_done[st], ph, abort_flag)) break;
      if (lane == 0) TR(3);
#pragma unroll
      for (int atom = 0; atom < 2; ++atom) {
        const uint8_t* src = smem + S::OFF_VR + vs * S::K_TILE + (atom * 64 + 2 * lane) * RBQ;
        uint8_t* vt = smem + S::OFF_VT + st * S::VT_TILE + atom * S::VT_ATOM + within;
#pragma unroll
        for (int v = 0; v < D / 8; ++v) {
          const uint4 x0 = *reinterpret_cast<const uint4*>(src + v * 16);            // key 2 lane
          const uint4 x1 = *reinterpret_cast<const uint4*>(src + RBQ + v * 16);      // key 2 lane + 1
          const uint32_t w0[4] = {x0.x, x0.y, x0.z, x0.w}, w1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int dd = v * 8 + e;                          // V^T row
            const uint32_t val = (e & 1) ? __byte_perm(w0[e >> 1], w1[e >> 1], 0x7632) : __byte_perm(w0[e >> 1], w1[e >> 1], 0x5410);
            *reinterpret_cast<uint32_t*>(vt + dd * 128 + ((chunk ^ (dd & 7)) << 4)) = val;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&vt_full[st]);
        mbar_arrive(&v_empty[vs]);
        TR(4);
      }
      if (++st == PS) { st = 0; ph ^= 1; }
    }
  } else {
    // ======================= softmax: thread = (query row, key half) =======================
    const int q = warp & 3, hf = (warp - 4) >> 2;
    const int row = q * 32 + lane;                          // TMEM lane = row of the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = t_lane + (uint32_t)(hf * 64);
    const uint32_t t_o = t_lane + (uint32_t)(S::O_COL + hf * S::NV);
    const bool warp_active = q0 + q * 32 < L;               // all 32 rows padding: keep the barriers moving, skip the math
    const float c = a.scale_log2;
    const float lazy = LAZY / c;                            // headroom in raw score units
    const int last_valid = L - (NT - 1) * BKEY - hf * 64;   // valid keys of this half in the last tile (may be <= 0)
    float m = 0.f;                                          // reference maximum (raw score units), set by the first batch
    uint8_t* const prow0 = smem + S::OFF_P + hf * S::P_ATOM + row * 128;
    const int rsw = row & 7;
    int st = 0;
    uint32_t pph = 1;
    TR_DECL(2 + hf)
    for (int j = 0; j < NT; ++j) {
      if (lane == 0 && q == 0) TR(11);
      if (!mbar_wait_parked(s_full, (uint32_t)(j & 1), abort_flag)) break;
      tc_fence_after();
      uint8_t* prow = prow0 + st * S::P_TILE;
      bool dead = false;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t r[32];
        if (warp_active) {
          tmem_ld32(t_s + (uint32_t)(sub * 32), r);
          tmem_ld_wait();
        }
        if (sub == 1) {                                      // S is in registers: the MMA may overwrite the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
          if (lane == 0 && q == 0) TR(13);
        }
        const int vh = j < NT - 1 ? 32 : min(max(last_valid - sub * 32, 0), 32);   // valid scores of this batch
        if (warp_active && vh > 0) {
          // batch maximum of this half row; the reference only moves when it is exceeded by more than the headroom
          float t0 = -INFINITY, t1 = -INFINITY;
          if (vh == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              t0 = fmaxf(t0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
              t1 = fmaxf(t1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < vh) t0 = fmaxf(t0, __uint_as_float(r[i]));
          }
          const float tm = fmaxf(t0, t1);
          if (j == 0 && sub == 0) {
            m = tm;
          } else if (__any_sync(0xffffffffu, tm > m + lazy)) {
            const float m_new = tm > m + lazy ? tm : m;
            const float f = ex2f((m - m_new) * c);           // 1 for the lanes that keep their reference
            m = m_new;
            if (j > 0) {
              // O_hf holds the sums of tiles 0 .. j-1: wait for the PV MMAs of tile j-1, then scale this warp's rows
              const int pst = st == 0 ? PS - 1 : st - 1;
              if (!mbar_wait_parked(&pv_done[pst], (uint32_t)(((j - 1) / PS) & 1), abort_flag)) { dead = true; break; }
              tc_fence_after();
#pragma unroll
              for (int cb = 0; cb < S::NV; cb += 16) {
                uint32_t o[16];
                tmem_ld16_nowait(t_o + cb, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                tmem_st16(t_o + cb, o);
              }
              tmem_st_wait();
              tc_fence_before();
            }
            if (sub == 1) rescale_p(prow, 0, rsw, f);       // the batch of this tile that was written against the old reference
          }
        }
        if (sub == 0) {
          if (j >= PS && !mbar_wait_parked(&pv_done[st], pph, abort_flag)) { dead = true; break; }   // P stage free
          if (lane == 0 && q == 0) TR(14);
        }
        if (warp_active) {
          const float mc = m * c;
          if (vh == 32) exp_store<false, POLY>(r, c, mc, 32, prow, sub * 4, rsw);
          else if (vh > 0) exp_store<true, 0>(r, c, mc, vh, prow, sub * 4, rsw);
          else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(prow + (((sub * 4 + ch) ^ rsw) << 4)) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      if (dead) break;
      fence_proxy_async();                                   // P stores (generic proxy) -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      if (lane == 0 && q == 0) TR(15);
      if (++st == PS) { st = 0; pph ^= 1; }
    }
    // ---- merge the two key halves: out = (O_0 f_0 + O_1 f_1) / (l_0 f_0 + l_1 f_1)
    if (hf == 1) smx[row] = m;
    asm volatile("bar.sync 1, %0;" ::"n"(NSM) : "memory");
    const int lst = (NT - 1) % PS;
    if (hf == 0 && mbar_wait_parked(&pv_done[lst], (uint32_t)(((NT - 1) / PS) & 1), abort_flag)) {   // warp-uniform
      tc_fence_after();
      uint32_t l0[16], l1[16];
      tmem_ld16_nowait(t_o + D, l0);                         // column D = sum of P = softmax denominator
      tmem_ld16_nowait(t_o + S::NV + D, l1);
      tmem_ld_wait();
      // a half without a single valid key (L <= 64) never set its reference: its sums are zero, any finite m works
      const float m1 = smx[row], mm = fmaxf(m, m1);
      const float f0 = ex2f((m - mm) * c), f1 = ex2f((m1 - mm) * c);
      const float inv = 1.0f / (__uint_as_float(l0[0]) * f0 + __uint_as_float(l1[0]) * f1);
      const float g0 = f0 * inv, g1 = f1 * inv;
      __half* op = a.out + ((size_t)(row0 + q0 + row) * a.E + h * D);
#pragma unroll
      for (int cb = 0; cb < D; cb += 16) {
        uint32_t o0[16], o1[16];
        tmem_ld16_nowait(t_o + cb, o0);
        tmem_ld16_nowait(t_o + S::NV + cb, o1);
        tmem_ld_wait();
        if (q0 + row < L) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o0[8 * i + e]) * g0 + __uint_as_float(o1[8 * i + e]) * g1;
            uint4 o;
            o.x = pack2(v[0], v[1]);
            o.y = pack2(v[2], v[3]);
            o.z = pack2(v[4], v[5]);
            o.w = pack2(v[6], v[7]);
            *reinterpret_cast<uint4*>(op + cb + 8 * i) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

template <int D, int POLY>
static int launch(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  using S = Smem<D>;
  if (!g_encode) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    CNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) {
      set_error("attention_tc05: cuTensorMapEncodeTiled unavailable");
      return CNB_ERR_UNSUPPORTED;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(f);
  }
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    CNB_CUDA(cudaFuncSetAttribute(attention_tc05_kernel<D, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  Args a;
  memset(&a, 0, sizeof(a));
  const cuuint64_t dims[2] = {(cuuint64_t)3 * E, (cuuint64_t)B * L};
  const cuuint64_t strides[1] = {(cuuint64_t)3 * E * 2};
  static_assert(BQ == BKEY, "Q and K tiles share one tensor map");
  const cuuint32_t box_q[2] = {(cuuint32_t)D, (cuuint32_t)BQ};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz = D == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_encode(&a.map_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box_q, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS)
    r = g_encode(&a.map_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box_q, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tc05: cuTensorMapEncodeTiled failed (%d): B=%d L=%d E=%d", (int)r, B, L, E);
    return CNB_ERR_CUDA;
  }
  a.out = reinterpret_cast<__half*>(out);
  a.L = L;
  a.E = E;
  a.q_tiles = ceil_div(L, BQ);
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)D);
  dim3 grid(a.q_tiles, heads, B);
  attention_tc05_kernel<D, POLY><<<grid, NUM_THREADS, S::TOTAL, st>>>(a);
  CNB_LAUNCH_CHECK();
  return CNB_OK;
}

}  // namespace atc5

int attention_tc05_error_flag() { return tc_read_clear_error(); }

bool attention_tc05_default() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CNB_ATTN_TC05");
    on = e ? atoi(e) : 0;
  }
  return on != 0;
}

bool attention_tc05_supported(const void* qkv, const void* out, int B, int L, int E, int heads) {
  if (heads <= 0 || E % heads) return false;
  const int d = E / heads;
  if (d != 16 && d != 32) return false;
  if (L < 96 || B > 65535 || heads > 65535) return false;
  if ((3 * E * 2) % 16 || (((uintptr_t)qkv | (uintptr_t)out) & 15)) return false;
  return true;
}

int attention_tc05(const void* qkv, void* out, int B, int L, int E, int heads, cudaStream_t st) {
  static int poly = -1;                                    // eighths of the exponentials evaluated on the FMA pipe
  if (poly < 0) {
    const char* e = getenv("CNB_ATTN_POLY");
    poly = e ? atoi(e) : 0;
  }
  const bool d16 = E / heads == 16;
  switch (poly) {
    case 1: return d16 ? atc5::launch<16, 1>(qkv, out, B, L, E, heads, st) : atc5::launch<32, 1>(qkv, out, B, L, E, heads, st);
    case 2: return d16 ? atc5::launch<16, 2>(qkv, out, B, L, E, heads, st) : atc5::launch<32, 2>(qkv, out, B, L, E, heads, st);
    case 3: return d16 ? atc5::launch<16, 3>(qkv, out, B, L, E, heads, st) : atc5::launch<32, 3>(qkv, out, B, L, E, heads, st);
    case 4: return d16 ? atc5::launch<16, 4>(qkv, out, B, L, E, heads, st) : atc5::launch<32, 4>(qkv, out, B, L, E, heads, st);
    default: return d16 ? atc5::launch<16, 0>(qkv, out, B, L, E, heads, st) : atc5::launch<32, 0>(qkv, out, B, L, E, heads, st);
  }
}

}  // namespace cnb
