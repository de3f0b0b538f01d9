// Fused  qkv projection -> octree window attention  on the 5th-gen tensor cores (tcgen05 / TMEM):
//
//   o = softmax( (y Wq^T + bq)(y Wk^T + bk)^T * scale + [same-submap mask] + RPE ) (y Wv^T + bv)
//
// for every (window, head) of a level -- OctreeAttention.forward up to (not including) `proj`
// (reference: models/octformer_backbone.py:52-88, RPE models/layers/octformer_layers.py:144-170,
// masks models/octree.py:186-209).  The (rows x 3C) qkv activation of the unfused schedule
// (3 KB per token and block at C = 256, written by the projection GEMM and read back by the attention
// kernel) never exists: a CTA owns a tile of TWO windows (128 rows: window slots of 64 rows) and, per
// group of four heads,
//   1. QKV chunk  D[128 x 192] = y_tile . Wg^T          tcgen05.mma M=128 N=192 (K = C), y tile by TMA,
//                                                        weights streamed from L2 through a TMA ring
//   2. drain      TMEM -> +bias -> bf16 -> shared memory as UMMA operands.  Per window: a BLOCK-DIAGONAL Q tile
//                 (row (parity, query) holds the 16 dims of head 2 hp + parity at K columns hp * 32 + parity * 16,
//                 zeros in the other parity's columns), a K tile (row key, same columns) and V^T tiles
//                 (row parity * 16 + dim, columns = keys) per head pair hp
//   3. per unit = (window, head pair):
//                 S[(parity, query)][key] = Q_{2 hp + parity}[query] . K_{2 hp + parity}[key]
//                                                        TWO tcgen05.mma (M=128, N=64, K=32) into a 64-column
//                                                        accumulator: every entry is a needed score
//                 softmax: thread == (query, head) row with ALL its keys (tcgen05.ld; row max / row sum are
//                 thread-local); bias = three fp16-pair table look-ups per (query, key) at offsets cached in
//                 registers for the whole tile, one look-up serving both units of the group; P (bf16) goes
//                 back INTO tensor memory over the scores (tcgen05.st, columns 0-31)
//                 O[(parity, query)][parity' * 16 + dim] = P . V_{2 hp + parity'}
//                                                        tcgen05.mma, A operand in tensor memory (M=128, N=32,
//                                                        K = keys rounded up to 16) into columns 32-63 of the
//                                                        unit's buffer; the row keeps the half of its own head
//   4. O drain    TMEM -> 1/rowsum -> bf16 -> global (32 B per row and head)
// Warp roles (320 threads, 168 registers, 1 CTA / SM, persistent over tiles):
//   0-7  softmax / drain: team = warp / 4 = window slot of the tile, lane quadrant = warp % 4; per head group
//        the team runs its two units TOGETHER (one bias look-up serves both heads of the thread), and while their
//        PV products run it stages the next group: Q / K as soon as every S product has completed, V^T once the
//        PV products have; the two teams only meet at the MMA barriers.  The (query, key) pair codes come from
//        k_pair_codes (made once per level, block-invariant) or are derived per tile when the caller passes none
//   8    tcgen05.mma issue + TMEM alloc          9   TMA producer (lanes 0-1 weights, lane 2 y tiles)
// Tensor memory (512 columns): QKV chunk 0-191 | four unit buffers of 64 columns at 192 + 64 (2 window + hp).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int QA_THREADS = 320;           // 8 softmax / drain warps + MMA issuer + TMA producer (168 registers)
constexpr int QA_SLOT_ROWS = 64;           // rows per window slot (2 slots per 128-row tile)
constexpr int QA_GN = 192;                 // QKV columns per head group: q | k | v of 4 heads
constexpr int QA_W_SLOT = QA_GN * 128;     // one K block of a group's weights: 192 rows x 64 bf16
constexpr int QA_RING = 3;
constexpr float QA_LOG2E = 1.4426950408889634f;

template <int C>
struct QaSmem {
  static constexpr int KB = C / 64;
  static constexpr int Y_BYTES = KB * 16384;
  static constexpr int OFF_W = Y_BYTES;
  static constexpr int OFF_A = OFF_W + QA_RING * QA_W_SLOT;   // [2 windows] 128 rows x 128 B: block-diagonal Q
  static constexpr int OFF_K = OFF_A + 2 * 16384;             // [2 windows] 64 keys x 128 B
  static constexpr int OFF_V = OFF_K + 2 * 8192;              // [2 windows][2 head pairs] V^T: 32 x 128 B (64 keys)
  static constexpr int OFF_TOK = OFF_V + 4 * 4096;            // [128] int4: 4x, 4y, 4z, submap
  static constexpr int OFF_BMAX = OFF_TOK + 2048;             // [16] fp32 largest bias of a head
  static constexpr int OFF_BIAS = OFF_BMAX + 64;              // [3C] fp32, group-major like the weights
  static constexpr int OFF_BAR = OFF_BIAS + 3 * C * 4;
  static constexpr int OFF_TAB = OFF_BAR + 256;               // RPE tables [3][H / 4][2][SUBP] fp16 pairs x log2 e
};

struct QaParams {
  __nv_bfloat16* out;          // [rows, C]
  const short4* xyzb;          // token table
  const float* rpe;            // [3 * (2 bnd + 1), H] or NULL
  const float* bias;           // [3C] group-major
  int n_win, H, K, dil, hat, bnd, subp;
  float scale;
  long long* prof;             // diagnostics only (HFL_QA_PROF): per-role wait / work cycles of CTA 0
  const uint32_t* codes;       // optional [n_win, L, LP] pair codes (hfl_qkv_attn_codes): block-invariant, made once per level
  int lp;                      // padded row length of `codes` (multiple of 4)
};

__device__ __forceinline__ void qa_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void qa_sts16(uint32_t a, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ float qa_lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t qa_lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t qa_hadd2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float qa_half_lo(uint32_t v) {
  return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu)));
}
__device__ __forceinline__ float qa_half_hi(uint32_t v) {
  return __half2float(__ushort_as_half((unsigned short)(v >> 16)));
}
// c + (fp16 half of a pair): mixed-precision add (FHADD), no separate conversion
__device__ __forceinline__ float qa_fhadd_lo(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tadd.rn.f32.f16 %0, l, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
__device__ __forceinline__ float qa_fhadd_hi(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tadd.rn.f32.f16 %0, h, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
__device__ __forceinline__ float qa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t qa_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// load / store N consecutive tensor-memory columns (N compile-time: powers of two, greedily)
template <int N>
__device__ __forceinline__ void qa_ld_cols(uint32_t a, uint32_t* v) {
  if constexpr (N >= 16) { uint32_t t[16]; ptx::tmem_ld16(a, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = t[i];
    qa_ld_cols<N - 16>(a + 16, v + 16); }
  else if constexpr (N >= 8) { ptx::tmem_ld8(a, v); qa_ld_cols<N - 8>(a + 8, v + 8); }
  else if constexpr (N >= 4) { ptx::tmem_ld4(a, v); qa_ld_cols<N - 4>(a + 4, v + 4); }
  else if constexpr (N >= 2) { ptx::tmem_ld2(a, v); qa_ld_cols<N - 2>(a + 2, v + 2); }
  else if constexpr (N >= 1) { ptx::tmem_ld1(a, v); }
}

#define QA_T0() const long long t0_ = PROF ? clock64() : 0
#define QA_ACC(slot) do { if (PROF) lacc[slot] += clock64() - t0_; } while (0)

// NKEY: keys of a window (K + relay token; <= 64) -- exact, so that the per-row register array of pair
// codes carries no padding
template <int C, int NKEY, bool PROF>
__global__ void __launch_bounds__(QA_THREADS, 1)
k_qkv_attn(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_w, const QaParams p) {
  using S = QaSmem<C>;
  constexpr int KB = S::KB, G = C / 64;                  // K blocks of the projection, head groups
  constexpr int NCH = (NKEY + 15) / 16;                  // 16-key chunks of a row = K steps of the PV product
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ptx::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sY = base, sW = base + S::OFF_W, sA = base + S::OFF_A, sK = base + S::OFF_K, sV = base + S::OFF_V;
  int4* s_tok = reinterpret_cast<int4*>(smem + S::OFF_TOK);
  float* s_bmax = reinterpret_cast<float*>(smem + S::OFF_BMAX);
  float* s_bias = reinterpret_cast<float*>(smem + S::OFF_BIAS);
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + S::OFF_TAB);
  const uint32_t bar = base + S::OFF_BAR;
  const uint32_t y_full = bar, y_empty = bar + 8;
  const uint32_t w_full = bar + 16, w_empty = bar + 16 + 8 * QA_RING;      // QA_RING each
  const uint32_t chunk_full = bar + 64, qk_ready = bar + 72, attn_done = bar + 80, v_ready = bar + 224;
  const uint32_t s_full = bar + 88;      // [4]  S of a unit is in tensor memory
  const uint32_t p_ready = bar + 120;    // [4]  P of a unit is in tensor memory
  const uint32_t o_full = bar + 152;     // [4]  O of a unit is in tensor memory
  const uint32_t o_free = bar + 184;     // [4]  O of a unit has been read: its buffer may take the next S
  const uint32_t s_tmem = bar + 232;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + S::OFF_BAR + 232);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, hat = p.hat, L = K + hat, H = p.H, dil = p.dil;
  const int tiles = (p.n_win + 1) >> 1;
  const int n_my = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int num = 2 * p.bnd + 1, SUBP = p.subp;
  long long lacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) lacc[i] = 0;
  const long long t_role = PROF ? clock64() : 0;

  // ---- one-time setup: bias vector, RPE tables (x log2 e; slot num = 0, slot num + 1 of the x axis = -inf),
  //      zeros in the block-diagonal Q tiles (the off-diagonal chunks are never written again) ----
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_bias[i] = p.bias[i];
  // one 4-byte entry = the fp16 biases of the two heads a softmax thread owns in a head group
  // (4 grp + par, 4 grp + 2 + par): one look-up per axis serves both units of the group
  for (int i = threadIdx.x; i < 3 * (H / 2) * SUBP; i += blockDim.x) {
    const int k = i % SUBP, gp = (i / SUBP) % (H / 2), axis = i / (SUBP * (H / 2));
    const int h0 = (gp >> 1) * 4 + (gp & 1), h1 = h0 + 2;
    float v0 = 0.f, v1 = 0.f;
    if (k < num) {
      if (p.rpe) {
        v0 = __ldg(p.rpe + (size_t)(axis * num + k) * H + h0) * QA_LOG2E;
        v1 = __ldg(p.rpe + (size_t)(axis * num + k) * H + h1) * QA_LOG2E;
      }
    } else if (k == num + 1 && axis == 0) {
      v0 = v1 = -INFINITY;
    }
    const __half2 hv = __floats2half2_rn(v0, v1);
    s_tab[i] = *reinterpret_cast<const uint32_t*>(&hv);
  }
  for (int i = threadIdx.x; i < 2 * 16384 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + S::OFF_A)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    ptx::mbar_init(y_full, 1);
    ptx::mbar_init(y_empty, 1);
    for (int s = 0; s < QA_RING; ++s) { ptx::mbar_init(w_full + 8 * s, 1); ptx::mbar_init(w_empty + 8 * s, 1); }
    ptx::mbar_init(chunk_full, 1);
    ptx::mbar_init(qk_ready, 8);
    ptx::mbar_init(v_ready, 8);
    ptx::mbar_init(attn_done, 1);
    for (int t = 0; t < 4; ++t) {
      ptx::mbar_init(s_full + 8 * t, 1);
      ptx::mbar_init(p_ready + 8 * t, 4);
      ptx::mbar_init(o_full + 8 * t, 1);
      ptx::mbar_init(o_free + 8 * t, 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 8) { ptx::tmem_alloc(s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();                              // the zeroed Q tiles are read by the tensor core
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // largest bias a head can add (sum of the per-axis table maxima): part of the exponent shift
  if (threadIdx.x < H) {
    const int h = threadIdx.x, gp = (h >> 2) * 2 + (h & 1), hi = (h >> 1) & 1;
    float tot = 0.f;
    for (int axis = 0; axis < 3; ++axis) {
      float m = 0.f;                                       // the zero slot is always reachable
      for (int k = 0; k < num; ++k) {
        const uint32_t e = s_tab[(axis * (H / 2) + gp) * SUBP + k];
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&e));
        m = fmaxf(m, hi ? f.y : f.x);
      }
      tot += m;
    }
    s_bmax[h] = tot;
  }
  __syncthreads();
  const uint32_t tmem_base = *tmem_ptr_s;
  // tensor memory: QKV chunk 0-191 | four unit buffers of 64 columns (S, then P in 0-31 and O in 32-63)
  constexpr uint32_t T_CHUNK = 0, T_U = 192;

  if (warp == 9) {
    // ===================== TMA producer =====================
    if (lane < 2) {
      ptx::prefetch_tmap(&tm_w);
      uint32_t g = 0;
      for (int it = 0; it < n_my; ++it)
        for (int grp = 0; grp < G; ++grp)
          for (int kb = 0; kb < KB; ++kb, ++g) {
            if ((int)(g & 1) != lane) continue;
            const uint32_t s = g % QA_RING, ph = (g / QA_RING) & 1;
            ptx::mbar_wait_sleep(w_empty + 8 * s, ph ^ 1, 64);
            ptx::mbar_arrive_expect_tx(w_full + 8 * s, QA_W_SLOT);
            ptx::tma_load_3d(sW + s * QA_W_SLOT, &tm_w, w_full + 8 * s, 0, grp * QA_GN, kb);
          }
    } else if (lane == 2) {
      ptx::prefetch_tmap(&tm_y);
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        ptx::mbar_wait_sleep(y_empty, (it & 1) ^ 1, 128);
        ptx::mbar_arrive_expect_tx(y_full, S::Y_BYTES);
        for (int ws = 0; ws < 2; ++ws) {
          const int w = tile * 2 + ws;
          // rows of window w: hat / plain: w * L + s; dilated: (w / dil) * K * dil + s * dil + w % dil
          // = row (w / dil) * K + s, phase w % dil of the {C, dil, rows / dil} view
          const int c1 = dil > 1 ? w % dil : 0;
          const int c2 = dil > 1 ? (w / dil) * K : w * L;
          for (int kb = 0; kb < KB; ++kb)
            ptx::tma_load_3d(sY + kb * 16384 + ws * 8192, &tm_y, y_full, kb * 64, c1, c2);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_qkv = ptx::umma_idesc_bf16(128, QA_GN);
      const uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64);
      const uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 32);
      uint32_t ring = 0;
      auto issue_qkv = [&](int grp) {
        for (int kb = 0; kb < KB; ++kb, ++ring) {
          const uint32_t s = ring % QA_RING, ph = (ring / QA_RING) & 1;
          { QA_T0(); ptx::mbar_wait(w_full + 8 * s, ph); QA_ACC(0); }
          ptx::tc_fence_after();
          const uint64_t ad = ptx::umma_desc_sw128(sY + kb * 16384);
          const uint64_t bd = ptx::umma_desc_sw128(sW + s * QA_W_SLOT);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem_base + T_CHUNK, ad + 2 * k, bd + 2 * k, idesc_qkv, (kb | k) != 0);
          ptx::umma_commit(w_empty + 8 * s);
        }
        ptx::umma_commit(chunk_full);
        if (grp == G - 1) ptx::umma_commit(y_empty);
      };
      // Flat loop over (tile, head group).  Per group: S of the four units as soon as Q / K are staged and the
      // units' buffers were emptied -> the next projection chunk as soon as V^T is staged too (the chunk columns
      // are free then) -> the PV products in the order the teams deliver P.
      const int total = n_my * G;
      { QA_T0(); ptx::mbar_wait(y_full, 0); QA_ACC(1); }
      ptx::tc_fence_after();
      issue_qkv(0);
      for (int gi = 0; gi < total; ++gi) {
        const int it = gi / G, grp = gi - it * G;
        const uint32_t ph = gi & 1;
        { QA_T0(); ptx::mbar_wait(qk_ready, ph); QA_ACC(2); }      // Q / K of this group are in shared memory
        // unit t = (window ws = t / 2, head pair hp = t % 2): S[(parity, query)][key] for both heads of the
        // pair in ONE accumulator through the block-diagonal Q tile (K = 32: 16 dims x 2 heads)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int ws = t >> 1, hp = t & 1;
          if (gi > 0) { QA_T0(); ptx::mbar_wait(o_free + 8 * t, ph ^ 1); QA_ACC(4); }
          ptx::tc_fence_after();
          const uint64_t ad = ptx::umma_desc_sw128(sA + ws * 16384) + 4 * hp;
          const uint64_t bd = ptx::umma_desc_sw128(sK + ws * 8192) + 4 * hp;
          ptx::umma_bf16(tmem_base + T_U + t * 64, ad, bd, idesc_s, 0);
          ptx::umma_bf16(tmem_base + T_U + t * 64, ad + 2, bd + 2, idesc_s, 1);
          ptx::umma_commit(s_full + 8 * t);
        }
        { QA_T0(); ptx::mbar_wait(v_ready, ph); QA_ACC(2); }       // V^T staged: the chunk columns are free
        ptx::tc_fence_after();
        // next projection chunk; across the tile boundary only if the next y tile has already landed (never
        // stall the PV issue on it), otherwise after the PV products
        bool deferred = false;
        if (gi + 1 < total) {
          if (grp + 1 < G) issue_qkv(grp + 1);
          else if (ptx::mbar_try_wait(y_full, (it + 1) & 1)) { ptx::tc_fence_after(); issue_qkv(0); }
          else deferred = true;
        }
        // PV products in the order the TEAMS deliver P (a team delivers both of its units together): whichever
        // window's softmax finishes first gets its products first, the other team is not made to wait behind it
        {
          QA_T0();
          uint32_t pend = 3u, spins = 0;
          while (pend) {
#pragma unroll
            for (int w2 = 0; w2 < 2; ++w2) {
              if (!((pend >> w2) & 1u)) continue;
              if (!ptx::mbar_try_wait(p_ready + 8 * (2 * w2), ph) || !ptx::mbar_try_wait(p_ready + 8 * (2 * w2 + 1), ph)) continue;
              ptx::tc_fence_after();
#pragma unroll
              for (int hp = 0; hp < 2; ++hp) {
                const int t = 2 * w2 + hp;
                const uint64_t vd = ptx::umma_desc_sw128(sV + t * 4096);
#pragma unroll
                for (int kk = 0; kk < NCH; ++kk)
                  ptx::umma_bf16_ts(tmem_base + T_U + t * 64 + 32, tmem_base + T_U + t * 64 + 8 * kk, vd + 2 * kk,
                                    idesc_pv, kk != 0);
                ptx::umma_commit(o_full + 8 * t);
              }
              pend &= ~(1u << w2);
            }
            if (++spins > (1u << 26)) __trap();
          }
          QA_ACC(3);
        }
        ptx::umma_commit(attn_done);
        if (deferred) {
          { QA_T0(); ptx::mbar_wait(y_full, (it + 1) & 1); QA_ACC(1); }
          ptx::tc_fence_after();
          issue_qkv(0);
        }
      }
      if (PROF) lacc[5] = clock64() - t_role;
    }
    __syncwarp();
  } else {
    // ===================== softmax / drain warps 0-7 =====================
    // team = warp / 4 owns window slot ws = team of the tile: per head group its two units (head pairs), one
    // after the other; quad = TMEM lane quadrant (hardware: warp % 4).  In a unit's accumulator lane
    // r = (head parity r / 64, query slot r % 64): the thread owns one (query, head) row with ALL its keys --
    // row maximum and row sum are thread-local, nothing is exchanged between threads.
    const int wq = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int quad = wq & 3, team = wq >> 2;
    const int r = quad * 32 + lane;
    const int par = r >> 6, sl = r & 63;                   // softmax view of the lane
    const int ws = team;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sc = p.scale * QA_LOG2E;
    const bool use_rpe = p.rpe != nullptr;
    const uint32_t tab_u = base + S::OFF_TAB;
    const int o_zero = num * 4, o_inf = (num + 1) * 4;
    const int bnd4 = p.bnd * 4;
    const int total = n_my * G;
    const int yw = r >> 6;                                 // drain view of the lane: row r of the y tile
    // +bias, bf16, UMMA operand layouts for one 16-column unit of the QKV chunk of group `grp` (u / 4 =
    // q | k | v, u % 4 = head hl of the group).  Here lane r is row r of the y tile (window yw, slot sl).
    auto drain_unit = [&](int grp, int sect, int hl, const uint32_t (&raw)[16]) {
      const int dpar = hl & 1, dhp = hl >> 1, u = sect * 4 + hl;
      const float* bg = s_bias + grp * QA_GN + u * 16;
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bg + 4 * q);
        v[4 * q] = __uint_as_float(raw[4 * q]) + b.x;
        v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
        v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
        v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
      }
      if (sect < 2) {
        // K-major 128B-swizzled tiles, 16-byte chunk ch of row rr at (ch ^ (rr & 7)).  Q: row (parity, slot) of
        // the window's block-diagonal tile; K: row slot.  The head's 16 dims are the chunks
        // 4 hp + 2 parity + {0, 1}: K columns hp * 32 + parity * 16 + d
        const int rr = sect == 0 ? dpar * 64 + sl : sl;
        const uint32_t rowa = (sect == 0 ? sA + (uint32_t)yw * 16384u : sK + (uint32_t)yw * 8192u) + (uint32_t)rr * 128u;
        const int ch = dhp * 4 + dpar * 2;
#pragma unroll
        for (int q = 0; q < 2; ++q)
          qa_sts128(rowa + (uint32_t)(((ch + q) ^ (rr & 7)) << 4), qa_pack(v[8 * q], v[8 * q + 1]),
                    qa_pack(v[8 * q + 2], v[8 * q + 3]), qa_pack(v[8 * q + 4], v[8 * q + 5]),
                    qa_pack(v[8 * q + 6], v[8 * q + 7]));
      } else {
        // V^T of the (window, head pair): row n = parity * 16 + d, 64 keys = 128 B
        const uint32_t keya = sV + (uint32_t)(yw * 2 + dhp) * 4096u + (uint32_t)(sl & 7) * 2u;
        const uint32_t kch = (uint32_t)sl >> 3;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          const __nv_bfloat16 hv = __float2bfloat16(v[d]);
          const uint32_t n = (uint32_t)(dpar * 16 + d);
          qa_sts16(keya + n * 128u + ((kch ^ (n & 7u)) << 4), *reinterpret_cast<const uint16_t*>(&hv));
        }
      }
    };
    // the team takes q, k and v of the heads 2 team, 2 team + 1 of the group
    auto drain_qk = [&](int gi) {
      const long long t_d = PROF ? clock64() : 0;
      const int grp = gi % G;
      // the next unit's tensor-memory load is in flight while the current one is converted and stored
      uint32_t raw[2][16];
      ptx::tmem_ld16(lane_base + T_CHUNK + (team * 2) * 16, raw[0]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ptx::tmem_ld_wait();
        if (i + 1 < 4) ptx::tmem_ld16(lane_base + T_CHUNK + (((i + 1) >> 1) * 4 + team * 2 + ((i + 1) & 1)) * 16, raw[(i + 1) & 1]);
        drain_unit(grp, i >> 1, team * 2 + (i & 1), raw[i & 1]);
      }
      ptx::fence_proxy_async();                            // generic-proxy stores -> UMMA (async proxy)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qk_ready);
      if (PROF) lacc[10] += clock64() - t_d;
    };
    auto drain_v = [&](int gi) {
      const long long t_d = PROF ? clock64() : 0;
      const int grp = gi % G;
      uint32_t raw[2][16];
      ptx::tmem_ld16(lane_base + T_CHUNK + (2 * 4 + team * 2) * 16, raw[0]);
      ptx::tmem_ld16(lane_base + T_CHUNK + (2 * 4 + team * 2 + 1) * 16, raw[1]);
      ptx::tmem_ld_wait();
      drain_unit(grp, 2, team * 2, raw[0]);
      drain_unit(grp, 2, team * 2 + 1, raw[1]);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(v_ready);
      if (PROF) lacc[10] += clock64() - t_d;
    };

    // prologue: stage group 0 of the first tile
    if (total > 0) {
      { QA_T0(); ptx::mbar_wait(chunk_full, 0); QA_ACC(8); }
      ptx::tc_fence_after();
      drain_qk(0);
      drain_v(0);
    }
    uint32_t code[NKEY];
    bool valid = false;
    int64_t row = 0;
    for (int gi = 0; gi < total; ++gi) {
      const int it = gi / G, grp = gi - it * G;
      const uint32_t gcount = (uint32_t)gi;
      if (grp == 0 && p.codes != nullptr) {
        // pair codes made once per level by k_pair_codes (they do not depend on the block): 16-byte loads of this
        // row's codes; both head parities of a query read the same line
        const int tile = blockIdx.x + it * gridDim.x;
        const long long t_codes = PROF ? clock64() : 0;
        const int w = tile * 2 + ws;
        valid = w < p.n_win && sl < L;
        if (hat) row = (int64_t)w * L + sl;
        else if (dil > 1) row = (int64_t)(w / dil) * K * dil + (int64_t)sl * dil + (w % dil);
        else row = (int64_t)w * K + sl;
        const uint4* cp = reinterpret_cast<const uint4*>(p.codes + ((size_t)(valid ? w : 0) * L + (valid ? sl : 0)) * p.lp);
#pragma unroll
        for (int j4 = 0; j4 < (NKEY + 3) / 4; ++j4) {
          const uint4 c = __ldg(cp + j4);
          if (4 * j4 < NKEY) code[4 * j4] = c.x;
          if (4 * j4 + 1 < NKEY) code[4 * j4 + 1] = c.y;
          if (4 * j4 + 2 < NKEY) code[4 * j4 + 2] = c.z;
          if (4 * j4 + 3 < NKEY) code[4 * j4 + 3] = c.w;
        }
        if (PROF) lacc[7] += clock64() - t_codes;
      } else if (grp == 0) {
        const int tile = blockIdx.x + it * gridDim.x;
        // token table of the tile: thread r of team 0 loads the token of y-tile row r (window r / 64, slot r % 64)
        if (team == 0) {
          const int w0 = tile * 2 + (r >> 6);
          const bool v0 = w0 < p.n_win && sl < L;
          int64_t tok0;
          if (hat) tok0 = (int64_t)w0 * K + (sl == 0 ? 0 : sl - 1);
          else if (dil > 1) tok0 = (int64_t)(w0 / dil) * K * dil + (int64_t)sl * dil + (w0 % dil);
          else tok0 = (int64_t)w0 * K + sl;
          // coordinates pre-multiplied by 4: clamp(4 dx, +-4 bnd) + 4 bnd is the byte offset into a table of 4-byte entries
          const short4 tk = v0 ? __ldg(p.xyzb + tok0) : make_short4(0, 0, 0, -2);
          s_tok[r] = make_int4(4 * (int)tk.x, 4 * (int)tk.y, 4 * (int)tk.z, (int)tk.w);
        }
        { QA_T0(); asm volatile("bar.sync 1, 256;" ::: "memory"); QA_ACC(6); }
        const long long t_codes = PROF ? clock64() : 0;
        const int w = tile * 2 + ws;
        valid = w < p.n_win && sl < L;
        if (hat) row = (int64_t)w * L + sl;
        else if (dil > 1) row = (int64_t)(w / dil) * K * dil + (int64_t)sl * dil + (w % dil);
        else row = (int64_t)w * K + sl;
        // ---- pair codes of the row (head-invariant, kept for the whole tile), branch-free:
        //      x | y << 10 | z << 20 byte offsets into the per-head-pair tables ----
        const int4 me = s_tok[ws * 64 + sl];
        const bool row_norel = !use_rpe || (hat && sl == 0);
#pragma unroll
        for (int j = 0; j < NKEY; ++j) {
          const int4 kj = s_tok[ws * 64 + j];
          const int ox = min(max(me.x - kj.x, -bnd4), bnd4) + bnd4;
          const int oy = min(max(me.y - kj.y, -bnd4), bnd4) + bnd4;
          const int oz = min(max(me.z - kj.z, -bnd4), bnd4) + bnd4;
          const bool same = me.w == kj.w;
          const bool norel = row_norel || (hat && j == 0);
          uint32_t a = (uint32_t)ox | ((uint32_t)oy << 10) | ((uint32_t)oz << 20);
          if (norel) a = (uint32_t)o_zero | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
          if (!same) a = (uint32_t)o_inf | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
          code[j] = a;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // s_tok may be rewritten for the next tile
        if (PROF) lacc[7] += clock64() - t_codes;

      }
      // ---- softmax of the team's two units TOGETHER: heads grp * 4 + par (unit 0) and grp * 4 + 2 + par
      //      (unit 1) of query (ws, sl); one table entry = the fp16 biases of exactly this pair of heads, so
      //      a (query, key) pair costs three look-ups for both units and nothing is cached ----
      const uint32_t tx = tab_u + (uint32_t)(((0 * G + grp) * 2 + par) * SUBP) * 4u;
      const uint32_t ty = tab_u + (uint32_t)(((1 * G + grp) * 2 + par) * SUBP) * 4u;
      const uint32_t tz = tab_u + (uint32_t)(((2 * G + grp) * 2 + par) * SUBP) * 4u;
      const uint32_t t_u0 = lane_base + T_U + (ws * 2) * 64, t_u1 = t_u0 + 64;
      float lsum[2];
      {
        { QA_T0(); ptx::mbar_wait(s_full + 8 * (ws * 2), gcount & 1); ptx::mbar_wait(s_full + 8 * (ws * 2 + 1), gcount & 1); QA_ACC(11); }
        ptx::tc_fence_after();
        const long long t_sm = PROF ? clock64() : 0;
        constexpr int NC8 = (NKEY + 7) / 8;                  // 8-key chunks of a row
        constexpr int NLAST = NKEY - (NC8 - 1) * 8;          // keys of the last chunk (1 .. 8)
        uint32_t raw[2][2][8];                               // [buffer][unit][key of the chunk]: the next chunk is in flight
        auto load_chunk = [&](int c, uint32_t (&dst)[2][8]) {
          if (c < NC8 - 1) { qa_ld_cols<8>(t_u0 + c * 8, dst[0]); qa_ld_cols<8>(t_u1 + c * 8, dst[1]); }
          else { qa_ld_cols<NLAST>(t_u0 + c * 8, dst[0]); qa_ld_cols<NLAST>(t_u1 + c * 8, dst[1]); }
        };
        // pass 1: an upper bound of the row maxima that needs no bias look-ups:
        // max_j(raw) * sc + (largest table sum of the head)
        float mx0 = -INFINITY, mx1 = -INFINITY;
        load_chunk(0, raw[0]);
#pragma unroll
        for (int c = 0; c < NC8; ++c) {
          ptx::tmem_ld_wait();
          if (c + 1 < NC8) load_chunk(c + 1, raw[(c + 1) & 1]);
          else load_chunk(0, raw[(c + 1) & 1]);              // first chunk of pass 2
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c * 8 + j < NKEY) {
              mx0 = fmaxf(mx0, __uint_as_float(raw[c & 1][0][j]));
              mx1 = fmaxf(mx1, __uint_as_float(raw[c & 1][1][j]));
            }
        }
        const int h0 = grp * 4 + par;
        // rows that do not exist (slots beyond the window, windows beyond the level) get p = 2^(-inf) = 0
        const float nshift0 = valid ? -fmaf(mx0, sc, s_bmax[h0]) : -INFINITY;
        const float nshift1 = valid ? -fmaf(mx1, sc, s_bmax[h0 + 2]) : -INFINITY;
        // pass 2: p = 2^(s * sc + bias - shift), bf16 pairs along the keys back over the scores
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < NC8; ++c) {
          const int bi = (NC8 + c) & 1;                      // buffer the chunk was loaded into
          ptx::tmem_ld_wait();
          if (c + 1 < NC8) load_chunk(c + 1, raw[bi ^ 1]);
          float pe0[8], pe1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pe0[j] = pe1[j] = 0.f;
            if (c * 8 + j < NKEY) {
              const uint32_t cd = code[c * 8 + j];
              const uint32_t bs = qa_hadd2(qa_hadd2(qa_lds_u32(tx + (cd & 0x3ffu)), qa_lds_u32(ty + ((cd >> 10) & 0x3ffu))),
                                           qa_lds_u32(tz + (cd >> 20)));
              pe0[j] = qa_ex2(qa_fhadd_lo(bs, fmaf(__uint_as_float(raw[bi][0][j]), sc, nshift0)));
              pe1[j] = qa_ex2(qa_fhadd_hi(bs, fmaf(__uint_as_float(raw[bi][1][j]), sc, nshift1)));
            }
          }
          uint32_t pk0[4], pk1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            l0 += pe0[2 * j] + pe0[2 * j + 1];
            l1 += pe1[2 * j] + pe1[2 * j + 1];
            pk0[j] = qa_pack(pe0[2 * j], pe0[2 * j + 1]);
            pk1[j] = qa_pack(pe1[2 * j], pe1[2 * j + 1]);
          }
          ptx::tmem_st4(t_u0 + c * 4, pk0);
          ptx::tmem_st4(t_u1 + c * 4, pk1);
        }
        if (NC8 & 1) {                                       // the PV product reads whole 16-key steps
          const uint32_t z[4] = {0u, 0u, 0u, 0u};
          ptx::tmem_st4(t_u0 + NC8 * 4, z);
          ptx::tmem_st4(t_u1 + NC8 * 4, z);
        }
        lsum[0] = l0;
        lsum[1] = l1;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(p_ready + 8 * (ws * 2)); ptx::mbar_arrive(p_ready + 8 * (ws * 2 + 1)); }
        if (PROF) lacc[12] += clock64() - t_sm;
      }

      // ---- while the PV products run: stage Q / K of the next group (S of this group has been read by every
      //      team's softmax only after ALL four S products completed -- they complete in issue order) ----
      if (gi + 1 < total) {
        { QA_T0(); ptx::mbar_wait(chunk_full, (gcount + 1) & 1); QA_ACC(8); }
        { QA_T0(); ptx::mbar_wait(s_full + 8 * 3, gcount & 1); QA_ACC(9); }
        ptx::tc_fence_after();
        drain_qk(gi + 1);
      }
      // ---- outputs of the two units: O / rowsum -> bf16 -> global (32 B per row and head) ----
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        const int t = ws * 2 + hp, h = grp * 4 + hp * 2 + par;
        const uint32_t t_u = lane_base + T_U + t * 64;
        { QA_T0(); ptx::mbar_wait(o_full + 8 * t, gcount & 1); QA_ACC(13); }
        ptx::tc_fence_after();
        const long long t_od = PROF ? clock64() : 0;
        uint32_t ro[16];
        ptx::tmem_ld16(t_u + 32 + par * 16, ro);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_free + 8 * t);
        if (valid) {
          const float s = 1.0f / fmaxf(lsum[hp], 1e-37f);
          uint4 a, b;
          a.x = qa_pack(__uint_as_float(ro[0]) * s, __uint_as_float(ro[1]) * s);
          a.y = qa_pack(__uint_as_float(ro[2]) * s, __uint_as_float(ro[3]) * s);
          a.z = qa_pack(__uint_as_float(ro[4]) * s, __uint_as_float(ro[5]) * s);
          a.w = qa_pack(__uint_as_float(ro[6]) * s, __uint_as_float(ro[7]) * s);
          b.x = qa_pack(__uint_as_float(ro[8]) * s, __uint_as_float(ro[9]) * s);
          b.y = qa_pack(__uint_as_float(ro[10]) * s, __uint_as_float(ro[11]) * s);
          b.z = qa_pack(__uint_as_float(ro[12]) * s, __uint_as_float(ro[13]) * s);
          b.w = qa_pack(__uint_as_float(ro[14]) * s, __uint_as_float(ro[15]) * s);
          uint4* dst = reinterpret_cast<uint4*>(p.out + row * C + h * 16);
          dst[0] = a;
          dst[1] = b;
        }
        if (PROF) lacc[14] += clock64() - t_od;
      }
      // ---- V^T of the next group once every PV product of this one has completed ----
      if (gi + 1 < total) {
        { QA_T0(); ptx::mbar_wait(attn_done, gcount & 1); QA_ACC(9); }
        ptx::tc_fence_after();
        drain_v(gi + 1);
      }
    }
    if (PROF) lacc[15] = clock64() - t_role;
  }
  if (PROF && blockIdx.x == 0) {
    if (threadIdx.x == 8 * 32) for (int i = 0; i < 6; ++i) p.prof[i] = lacc[i];
    if (threadIdx.x == 0) for (int i = 6; i < 16; ++i) p.prof[i] = lacc[i];
    if (threadIdx.x == 4 * 32) for (int i = 6; i < 16; ++i) p.prof[16 + i] = lacc[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// Pair codes of a level, once for all its blocks: codes[w][s][j] = x | y << 10 | z << 20 byte offsets of the
// (query s, key j) pair of window w into the RPE tables of k_qkv_attn (slot num = zero bias: no RPE for the relay
// token; slot num + 1 of the x axis = -inf: different submaps).  Rows are padded to lp = 4 ceil(L / 4) entries.
__global__ void __launch_bounds__(256) k_pair_codes(const short4* __restrict__ xyzb, uint32_t* __restrict__ codes,
                                                    int n_win, int K, int hat, int dil, int bnd, int use_rpe, int lp) {
  __shared__ int4 tk[64];
  const int w = blockIdx.x, L = K + hat;
  const int num = 2 * bnd + 1, o_zero = num * 4, o_inf = (num + 1) * 4, bnd4 = bnd * 4;
  for (int s = threadIdx.x; s < L; s += blockDim.x) {
    int64_t t;
    if (hat) t = (int64_t)w * K + (s == 0 ? 0 : s - 1);
    else if (dil > 1) t = (int64_t)(w / dil) * K * dil + (int64_t)s * dil + (w % dil);
    else t = (int64_t)w * K + s;
    const short4 v = __ldg(xyzb + t);
    tk[s] = make_int4(4 * (int)v.x, 4 * (int)v.y, 4 * (int)v.z, (int)v.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * lp; i += blockDim.x) {
    const int s = i / lp, j = i - s * lp;
    uint32_t a = 0;
    if (j < L) {
      const int4 me = tk[s], kj = tk[j];
      const int ox = min(max(me.x - kj.x, -bnd4), bnd4) + bnd4;
      const int oy = min(max(me.y - kj.y, -bnd4), bnd4) + bnd4;
      const int oz = min(max(me.z - kj.z, -bnd4), bnd4) + bnd4;
      const bool norel = !use_rpe || (hat && (s == 0 || j == 0));
      a = (uint32_t)ox | ((uint32_t)oy << 10) | ((uint32_t)oz << 20);
      if (norel) a = (uint32_t)o_zero | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
      if (me.w != kj.w) a = (uint32_t)o_inf | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
    }
    codes[((size_t)w * L + s) * lp + j] = a;
  }
}

typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);
static PFN_encodeTiled3 qa_get_encode() {
  static PFN_encodeTiled3 fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled3)q;
  }
  return fn;
}

template <int C, int NKEY, bool PROF>
static int launch_qa2(const CUtensorMap& ty, const CUtensorMap& tw, const QaParams& p, int smem, cudaStream_t st) {
  HFL_CUDA(cudaFuncSetAttribute(k_qkv_attn<C, NKEY, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int tiles = (p.n_win + 1) / 2, sms = sm_count();
  HFL_LAUNCH((k_qkv_attn<C, NKEY, PROF><<<tiles < sms ? tiles : sms, QA_THREADS, smem, st>>>(ty, tw, p)));
  return HFL_OK;
}
template <int C, int NKEY>
static int launch_qa(const CUtensorMap& ty, const CUtensorMap& tw, QaParams& p, int smem, cudaStream_t st) {
  static const bool want_prof = getenv("HFL_QA_PROF") != nullptr;
  if (!(want_prof && C == 256 && NKEY == 49)) return launch_qa2<C, NKEY, false>(ty, tw, p, smem, st);
  if constexpr (C == 256 && NKEY == 49) {
    static long long* prof = nullptr;
    if (!prof) cudaMalloc(&prof, 32 * sizeof(long long));
    cudaMemsetAsync(prof, 0, 32 * sizeof(long long), st);
    p.prof = prof;
    const int rc = launch_qa2<C, NKEY, true>(ty, tw, p, smem, st);
    long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* nm[] = {"mma:w_full", "mma:y_full", "mma:qkv_ready", "mma:p_ready", "mma:o_free", "mma:total",
                               "bar_tok", "codes", "chunk_full", "attn_done", "drain", "s_full", "softmax", "o_full",
                               "o_drain", "total"};
    const int tiles = (p.n_win + 1) / 2, sms = sm_count();
    fprintf(stderr, "[hfl_qkv_attn prof n_win=%d tiles/CTA=%.1f]", p.n_win, (double)tiles / (tiles < sms ? tiles : sms));
    for (int i = 0; i < 6; ++i) fprintf(stderr, " %s=%.1fk", nm[i], h[i] / 1e3);
    for (int i = 6; i < 16; ++i) fprintf(stderr, " t0:%s=%.1fk", nm[i], h[i] / 1e3);
    for (int i = 6; i < 16; ++i) fprintf(stderr, " t1:%s=%.1fk", nm[i], h[16 + i] / 1e3);
    fprintf(stderr, "\n");
    return rc;
  }
  return HFL_OK;
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_qkv_attn_supported(int32_t H, int32_t C, int32_t K, int32_t dil, int32_t hat, int32_t bnd) {
  const int L = K + (hat ? 1 : 0);
  if (!(C == 128 || C == 256) || C != H * 16) return 0;
  if (dil < 1 || (hat && dil != 1)) return 0;
  if (!(L == 64 || L == 49 || L == 48 || L == 33 || L == 32 || L == 17 || L == 16)) return 0;
  if ((2 * bnd + 3) > 256) return 0;
  const int subp = (2 * bnd + 3 + 3) & ~3;
  const int smem = (C == 128 ? QaSmem<128>::OFF_TAB : QaSmem<256>::OFF_TAB) + 3 * (H / 2) * subp * 4;
  return smem <= 227 * 1024;
}

int64_t hfl_qkv_attn_codes_bytes(int64_t n_win, int32_t K, int32_t hat) {
  const int L = K + (hat ? 1 : 0);
  return n_win * L * (int64_t)((L + 3) & ~3) * 4;
}

int hfl_qkv_attn_codes(const int16_t* xyzb, int64_t n_win, int32_t K, int32_t dil, int32_t hat, int32_t bnd,
                       int32_t use_rpe, uint32_t* codes, void* stream_) {
  if (n_win == 0) return HFL_OK;
  const int L = K + (hat ? 1 : 0);
  HFL_CHECK_ARG(xyzb && codes, "null argument");
  HFL_CHECK_ARG(L <= 64 && dil >= 1 && (!hat || dil == 1) && n_win % dil == 0 && (2 * bnd + 3) <= 256, "bad window shape");
  HFL_LAUNCH((k_pair_codes<<<(unsigned)n_win, 256, 0, (cudaStream_t)stream_>>>((const short4*)xyzb, codes, (int)n_win, K,
                                                                            hat ? 1 : 0, dil, bnd, use_rpe, (L + 3) & ~3)));
  return HFL_OK;
}

int hfl_qkv_attn(const void* y, const void* Wg, const float* bias_g, void* out, const int16_t* xyzb,
                 const float* rpe, int64_t n_win, int64_t rows, int32_t H, int32_t C, int32_t K, int32_t dil,
                 int32_t hat, int32_t bnd, float scale, const uint32_t* codes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_win == 0) return HFL_OK;
  HFL_CHECK_ARG(y && Wg && bias_g && out && xyzb, "null argument");
  HFL_CHECK_ARG(hfl_qkv_attn_supported(H, C, K, dil, hat, bnd), "configuration not supported by the fused qkv + attention kernel");
  HFL_CHECK_ARG(n_win % dil == 0 && rows % dil == 0 && n_win < (1ll << 30) && rows < (1ll << 31), "bad window / row count");
  PFN_encodeTiled3 enc = qa_get_encode();
  if (!enc) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled unavailable%s", "");
  CUtensorMap ty, tw;
  {
    // y as {C, dil, rows / dil}: a window's rows are 64 consecutive entries of the last axis at a fixed phase
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)dil, (cuuint64_t)(rows / dil)};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * dil};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)QA_SLOT_ROWS}, es[3] = {1, 1, 1};
    CUresult cr = enc(&ty, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(y), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (y) failed%s (%lld)", "", (long long)cr);
  }
  {
    // group-major weights [3C, C] as {64 K-columns, 3C rows, C / 64 K blocks}: one box = one K block of a group
    cuuint64_t dims[3] = {64, (cuuint64_t)(3 * C), (cuuint64_t)(C / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)QA_GN, 1}, es[3] = {1, 1, 1};
    CUresult cr = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(Wg), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (Wqkv) failed%s (%lld)", "", (long long)cr);
  }
  QaParams p;
  p.out = (__nv_bfloat16*)out; p.xyzb = (const short4*)xyzb; p.rpe = rpe; p.bias = bias_g;
  p.n_win = (int)n_win; p.H = H; p.K = K; p.dil = dil; p.hat = hat ? 1 : 0; p.bnd = bnd;
  p.subp = (2 * bnd + 3 + 3) & ~3;
  p.scale = scale;
  p.prof = nullptr;
  p.codes = codes;
  p.lp = (K + (hat ? 1 : 0) + 3) & ~3;
  const int nkey = K + p.hat;
  const int smem = (C == 128 ? QaSmem<128>::OFF_TAB : QaSmem<256>::OFF_TAB) + 3 * (H / 2) * p.subp * 4;
#define HFL_QA_CASE(C_, NK_) \
  if (C == C_ && nkey == NK_) return launch_qa<C_, NK_>(ty, tw, p, smem, st);
  HFL_QA_CASE(256, 64) HFL_QA_CASE(256, 49) HFL_QA_CASE(256, 48) HFL_QA_CASE(256, 33) HFL_QA_CASE(256, 32)
  HFL_QA_CASE(256, 17) HFL_QA_CASE(256, 16)
  HFL_QA_CASE(128, 64) HFL_QA_CASE(128, 49) HFL_QA_CASE(128, 48) HFL_QA_CASE(128, 33) HFL_QA_CASE(128, 32)
  HFL_QA_CASE(128, 17) HFL_QA_CASE(128, 16)
#undef HFL_QA_CASE
  return fail(HFL_ERR_UNSUPPORTED, "unsupported window size%s (%lld)", "", (long long)K);
}

}  // extern "C"
