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
//   2. drain      TMEM -> +bias -> bf16 -> shared memory as UMMA operands: Q, K (K-major, 128B swizzle)
//                 and V^T (dims x keys, K-major)
//   3. per head   S = Q_h K_h^T                          ONE tcgen05.mma (M=128, N=128, K=16) into TMEM
//                 softmax: thread == query row; the row's scores come out of tensor memory
//                 (tcgen05.ld), bias = three fp32 table look-ups per pair at offsets cached in registers
//                 for the whole tile (block-invariant pair codes), P (bf16) goes back INTO tensor memory
//                 over S (tcgen05.st)
//                 O_h = P V_h                            tcgen05.mma with the A operand in tensor memory
//                                                        (M=128, N=16, K=128: own window's keys, zeros for
//                                                        the other window slot)
//   4. O drain    TMEM -> 1/rowsum -> bf16 -> global (32 B per row and head)
// Warp roles (576 threads, 1 CTA / SM, persistent over tiles):
//   0-15 softmax / drain: lane quadrant = warp % 4 (thread == row), warp / 4 = which quarter of the row's
//        keys the thread evaluates (row maximum bound and row sum are combined through shared memory);
//        they also drain the QKV chunk (three 16-column units each) and one head's output each
//   16   tcgen05.mma issue + TMEM alloc          17  TMA producer (lanes 0-1 weights, lane 2 y tiles)
// Tensor memory (512 columns): QKV chunk 0-191 | S/P buffer 0: 192-319 | S/P buffer 1: 320-447 | O: 448-511.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int QA_THREADS = 576;           // 16 softmax / drain warps + MMA issuer + TMA producer
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
  static constexpr int OFF_Q = OFF_W + QA_RING * QA_W_SLOT;
  static constexpr int OFF_K = OFF_Q + 16384;
  static constexpr int OFF_V = OFF_K + 16384;          // V^T: [4 heads][2 key blocks][16 dims x 128 B]
  static constexpr int OFF_TOK = OFF_V + 16384;        // [128] int4: 4x, 4y, 4z, submap
  static constexpr int OFF_MX = OFF_TOK + 2048;        // [2][4 parts][128 rows] fp32 row-max exchange
  static constexpr int OFF_L = OFF_MX + 4096;          // [4 heads][4 parts][128 rows] fp32 partial row sums
  static constexpr int OFF_BMAX = OFF_L + 8192;        // [16] fp32 largest bias of a head
  static constexpr int OFF_BIAS = OFF_BMAX + 64;       // [3C] fp32, group-major like the weights
  static constexpr int OFF_BAR = OFF_BIAS + 3 * C * 4;
  static constexpr int OFF_TAB = OFF_BAR + 256;        // RPE tables [3][H / 2][SUBP] fp16 pairs (even, odd head) x log2 e
};

struct QaParams {
  __nv_bfloat16* out;          // [rows, C]
  const short4* xyzb;          // token table
  const float* rpe;            // [3 * (2 bnd + 1), H] or NULL
  const float* bias;           // [3C] group-major
  int n_win, H, K, dil, hat, bnd, subp;
  float scale;
  long long* prof;             // diagnostics only (HFL_QA_PROF): per-role wait / work cycles of CTA 0
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
__device__ __forceinline__ float qa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// store N consecutive tensor-memory columns from registers (N compile-time: powers of two, greedily)
template <int N>
__device__ __forceinline__ void qa_st_cols(uint32_t a, const uint32_t* v) {
  if constexpr (N >= 16) { uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = v[i];
    ptx::tmem_st16(a, t); qa_st_cols<N - 16>(a + 16, v + 16); }
  else if constexpr (N >= 8) { ptx::tmem_st8(a, v); qa_st_cols<N - 8>(a + 8, v + 8); }
  else if constexpr (N >= 4) { ptx::tmem_st4(a, v); qa_st_cols<N - 4>(a + 4, v + 4); }
  else if constexpr (N >= 2) { ptx::tmem_st2(a, v); qa_st_cols<N - 2>(a + 2, v + 2); }
  else if constexpr (N >= 1) { ptx::tmem_st1(a, v); }
}
// NV value columns followed by zeros up to NTOT columns
template <int NV, int NTOT>
__device__ __forceinline__ void qa_st_cols_tail(uint32_t a, const uint32_t* v) {
  uint32_t t[NTOT];
#pragma unroll
  for (int i = 0; i < NTOT; ++i) t[i] = i < NV ? v[i] : 0u;
  qa_st_cols<NTOT>(a, t);
}
__device__ __forceinline__ uint32_t qa_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// NKEY: keys of a window (K + relay token; <= 64) -- exact, so that the per-row register arrays (pair
// codes + scores) carry no padding
#define QA_T0() const long long t0_ = PROF ? clock64() : 0
#define QA_ACC(slot) do { if (PROF) lacc[slot] += clock64() - t0_; } while (0)

template <int C, int NKEY, bool PROF>
__global__ void __launch_bounds__(QA_THREADS, 1)
k_qkv_attn(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_w, const QaParams p) {
  using S = QaSmem<C>;
  constexpr int KB = S::KB, G = C / 64;                  // K blocks of the projection, head groups
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ptx::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sY = base, sW = base + S::OFF_W, sQ = base + S::OFF_Q, sK = base + S::OFF_K, sV = base + S::OFF_V;
  int4* s_tok = reinterpret_cast<int4*>(smem + S::OFF_TOK);
  float* s_mx = reinterpret_cast<float*>(smem + S::OFF_MX);
  float* s_l = reinterpret_cast<float*>(smem + S::OFF_L);
  float* s_bmax = reinterpret_cast<float*>(smem + S::OFF_BMAX);
  float* s_bias = reinterpret_cast<float*>(smem + S::OFF_BIAS);
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + S::OFF_TAB);
  const uint32_t bar = base + S::OFF_BAR;
  const uint32_t y_full = bar, y_empty = bar + 8;
  const uint32_t w_full = bar + 16, w_empty = bar + 16 + 8 * QA_RING;      // QA_RING each
  const uint32_t chunk_full = bar + 64, qkv_ready = bar + 72, attn_done = bar + 80;
  const uint32_t s_full = bar + 88;      // [2]
  const uint32_t p_ready = bar + 104;    // [2]
  const uint32_t o_full = bar + 120;     // [4]
  const uint32_t o_free = bar + 152;     // [4]
  const uint32_t s_tmem = bar + 184;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + S::OFF_BAR + 184);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, hat = p.hat, L = K + hat, H = p.H, dil = p.dil;
  const int tiles = (p.n_win + 1) >> 1;
  const int n_my = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int num = 2 * p.bnd + 1, SUBP = p.subp;
  long long lacc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) lacc[i] = 0;
  const long long t_role = PROF ? clock64() : 0;

  // ---- one-time setup: bias vector, RPE tables (x log2 e; slot num = 0, slot num + 1 of the x axis = -inf) ----
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_bias[i] = p.bias[i];
  // one 4-byte entry = the biases of an (even, odd) pair of heads as fp16: a look-up serves two heads and
  // neighbouring offsets share a bank word (half the bank conflicts of an fp32 table)
  for (int i = threadIdx.x; i < 3 * (H / 2) * SUBP; i += blockDim.x) {
    const int k = i % SUBP, hp = (i / SUBP) % (H / 2), axis = i / (SUBP * (H / 2));
    float v0 = 0.f, v1 = 0.f;
    if (k < num) {
      if (p.rpe) {
        v0 = __ldg(p.rpe + (size_t)(axis * num + k) * H + 2 * hp) * QA_LOG2E;
        v1 = __ldg(p.rpe + (size_t)(axis * num + k) * H + 2 * hp + 1) * QA_LOG2E;
      }
    } else if (k == num + 1 && axis == 0) {
      v0 = v1 = -INFINITY;
    }
    const __half2 hv = __floats2half2_rn(v0, v1);
    s_tab[i] = *reinterpret_cast<const uint32_t*>(&hv);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(y_full, 1);
    ptx::mbar_init(y_empty, 1);
    for (int s = 0; s < QA_RING; ++s) { ptx::mbar_init(w_full + 8 * s, 1); ptx::mbar_init(w_empty + 8 * s, 1); }
    ptx::mbar_init(chunk_full, 1);
    ptx::mbar_init(qkv_ready, 16);
    ptx::mbar_init(attn_done, 1);
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(s_full + 8 * b, 1); ptx::mbar_init(p_ready + 8 * b, 16); }
    for (int h = 0; h < 4; ++h) { ptx::mbar_init(o_full + 8 * h, 1); ptx::mbar_init(o_free + 8 * h, 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 16) { ptx::tmem_alloc(s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // largest bias a head can add (sum of the per-axis table maxima): part of the exponent shift
  if (threadIdx.x < H) {
    float tot = 0.f;
    for (int axis = 0; axis < 3; ++axis) {
      float m = 0.f;                                       // the zero slot is always reachable
      for (int k = 0; k < num; ++k) {
        const uint32_t e = s_tab[(axis * (H / 2) + (threadIdx.x >> 1)) * SUBP + k];
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&e));
        m = fmaxf(m, (threadIdx.x & 1) ? f.y : f.x);
      }
      tot += m;
    }
    s_bmax[threadIdx.x] = tot;
  }
  __syncthreads();
  const uint32_t tmem_base = *tmem_ptr_s;
  constexpr uint32_t T_CHUNK = 0, T_S = 192, T_O = 448;

  if (warp == 17) {
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
  } else if (warp == 16) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_qkv = ptx::umma_idesc_bf16(128, QA_GN);
      const uint32_t idesc_s = ptx::umma_idesc_bf16(128, 128);
      const uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 16);
      uint32_t ring = 0, gcount = 0, pcnt[2] = {0, 0};
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
      auto issue_s = [&](int hl) {
        const uint32_t buf = hl & 1;
        ptx::umma_bf16(tmem_base + T_S + buf * 128, ptx::umma_desc_sw128(sQ) + 2 * hl,
                       ptx::umma_desc_sw128(sK) + 2 * hl, idesc_s, 0);
        ptx::umma_commit(s_full + 8 * buf);
      };
      for (int it = 0; it < n_my; ++it) {
        { QA_T0(); ptx::mbar_wait(y_full, it & 1); QA_ACC(1); }
        ptx::tc_fence_after();
        issue_qkv(0);
        for (int grp = 0; grp < G; ++grp, ++gcount) {
          { QA_T0(); ptx::mbar_wait(qkv_ready, gcount & 1); QA_ACC(2); }   // Q / K / V^T of this group are in shared memory
          ptx::tc_fence_after();
          issue_s(0);
          issue_s(1);
          if (grp + 1 < G) issue_qkv(grp + 1);            // the chunk columns were drained before qkv_ready
          for (int hl = 0; hl < 4; ++hl) {
            const uint32_t buf = hl & 1;
            { QA_T0(); ptx::mbar_wait(p_ready + 8 * buf, pcnt[buf] & 1); QA_ACC(3); }
            ++pcnt[buf];
            if (gcount > 0) { QA_T0(); ptx::mbar_wait(o_free + 8 * hl, (gcount - 1) & 1); QA_ACC(4); }
            ptx::tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              ptx::umma_bf16_ts(tmem_base + T_O + hl * 16, tmem_base + T_S + buf * 128 + 8 * kk,
                                ptx::umma_desc_sw128(sV + hl * 4096 + (kk >> 2) * 2048) + 2 * (kk & 3), idesc_pv,
                                kk != 0);
            ptx::umma_commit(o_full + 8 * hl);
            if (hl + 2 < 4) issue_s(hl + 2);
          }
          ptx::umma_commit(attn_done);
        }
      }
      if (PROF) lacc[5] = clock64() - t_role;
    }
    __syncwarp();
  } else {
    // ===================== softmax / drain warps 0-15 =====================
    // quad = TMEM lane quadrant (hardware: warp % 4); the four warps of a quadrant split the KEYS of a row:
    // thread (row r, part) owns keys [part * KP, ...) of every head -- 13 keys instead of 49 per thread
    // keeps the per-thread state small enough for 16 resident softmax warps (4 per scheduler).
    const int wq = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int quad = wq & 3, part = wq >> 2;
    const int r = quad * 32 + lane, ws = r >> 6, sl = r & 63;
    constexpr int KP = (NKEY / 4) & ~1;                    // keys per part (even: bf16 pairs stay thread-local)
    constexpr int KMAX = NKEY - 3 * KP;                    // the last part takes the remainder (>= KP)
    static_assert(KMAX <= 16 && KP >= 2, "key split does not fit the x16 tensor-memory loads");
    const int k0 = part * KP, nk = part == 3 ? KMAX : KP;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sc = p.scale * QA_LOG2E;
    const bool use_rpe = p.rpe != nullptr;
    const uint32_t tab_u = base + S::OFF_TAB;
    const int o_zero = num * 4, o_inf = (num + 1) * 4;
    const int bnd4 = p.bnd * 4;
    uint32_t gcount = 0, hcount = 0;                       // groups done, heads done

    // output of head `part` of group number gc (group index grp_o within its tile): O / rowsum -> bf16 -> global
    auto drain_o = [&](uint32_t gc, int grp_o, int64_t row_o, bool valid_o) {
      const int hl = part, h = grp_o * 4 + hl;
      { QA_T0(); ptx::mbar_wait(o_full + 8 * hl, gc & 1); QA_ACC(13); }
      ptx::tc_fence_after();
      const long long t_od = PROF ? clock64() : 0;
      uint32_t ro[16];
      ptx::tmem_ld16(lane_base + T_O + hl * 16, ro);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_free + 8 * hl);
      if (valid_o) {
        const float* lp = s_l + hl * 512 + r;
        const float s = 1.0f / fmaxf((lp[0] + lp[128]) + (lp[256] + lp[384]), 1e-37f);
        uint4 a, b;
        a.x = qa_pack(__uint_as_float(ro[0]) * s, __uint_as_float(ro[1]) * s);
        a.y = qa_pack(__uint_as_float(ro[2]) * s, __uint_as_float(ro[3]) * s);
        a.z = qa_pack(__uint_as_float(ro[4]) * s, __uint_as_float(ro[5]) * s);
        a.w = qa_pack(__uint_as_float(ro[6]) * s, __uint_as_float(ro[7]) * s);
        b.x = qa_pack(__uint_as_float(ro[8]) * s, __uint_as_float(ro[9]) * s);
        b.y = qa_pack(__uint_as_float(ro[10]) * s, __uint_as_float(ro[11]) * s);
        b.z = qa_pack(__uint_as_float(ro[12]) * s, __uint_as_float(ro[13]) * s);
        b.w = qa_pack(__uint_as_float(ro[14]) * s, __uint_as_float(ro[15]) * s);
        uint4* dst = reinterpret_cast<uint4*>(p.out + row_o * C + h * 16);
        dst[0] = a;
        dst[1] = b;
      }
      if (PROF) lacc[14] += clock64() - t_od;
    };
    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int w = tile * 2 + ws;
      const bool valid = w < p.n_win && sl < L;
      // layout row of this thread's slot and the token behind it
      int64_t row, tok;
      if (hat) { row = (int64_t)w * L + sl; tok = (int64_t)w * K + (sl == 0 ? 0 : sl - 1); }
      else if (dil > 1) { row = (int64_t)(w / dil) * K * dil + (int64_t)sl * dil + (w % dil); tok = row; }
      else { row = (int64_t)w * K + sl; tok = row; }
      if (part == 0) {
        // coordinates pre-multiplied by 4: clamp(4 dx, +-4 bnd) + 4 bnd is the byte offset into an fp32 table
        const short4 tk = valid ? __ldg(p.xyzb + tok) : make_short4(0, 0, 0, -2);
        s_tok[r] = make_int4(4 * (int)tk.x, 4 * (int)tk.y, 4 * (int)tk.z, (int)tk.w);
      }
      { QA_T0(); asm volatile("bar.sync 1, 512;" ::: "memory"); QA_ACC(6); }
      const long long t_codes = PROF ? clock64() : 0;
      // ---- pair codes of this thread's keys (block-invariant within the tile), branch-free:
      //      x | y << 10 | z << 20 byte offsets into the per-head-pair tables ----
      const int4 me = s_tok[r];
      const bool row_norel = !use_rpe || (hat && sl == 0);
      uint32_t code[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int key = k0 + j;                            // < 64 always
        const int4 kj = s_tok[ws * 64 + key];
        const int ox = min(max(me.x - kj.x, -bnd4), bnd4) + bnd4;
        const int oy = min(max(me.y - kj.y, -bnd4), bnd4) + bnd4;
        const int oz = min(max(me.z - kj.z, -bnd4), bnd4) + bnd4;
        const bool same = me.w == kj.w;
        const bool norel = row_norel || (hat && key == 0);
        uint32_t a = (uint32_t)ox | ((uint32_t)oy << 10) | ((uint32_t)oz << 20);
        if (norel) a = (uint32_t)o_zero | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
        if (!same) a = (uint32_t)o_inf | ((uint32_t)o_zero << 10) | ((uint32_t)o_zero << 20);
        code[j] = a;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");      // s_tok may be rewritten for the next tile
      if (PROF) lacc[7] += clock64() - t_codes;

      for (int grp = 0; grp < G; ++grp, ++gcount) {
        // ---- drain the QKV chunk: +bias, bf16, UMMA operand layouts; 12 units of 16 columns
        //      (unit u: u / 4 = q | k | v, u % 4 = head of the group); part p takes q, k, v of head p ----
        { QA_T0(); ptx::mbar_wait(chunk_full, gcount & 1); QA_ACC(8); }
        if (gcount > 0) { QA_T0(); ptx::mbar_wait(attn_done, (gcount - 1) & 1); QA_ACC(9); }      // staging buffers free
        ptx::tc_fence_after();
        const long long t_drain2 = PROF ? clock64() : 0;
#pragma unroll 1
        for (int uu = 0; uu < 3; ++uu) {
          const int u = uu * 4 + part, sect = uu, hl = part;   // part p: q, k and v of head p (balanced)
          uint32_t raw[16];
          ptx::tmem_ld16(lane_base + T_CHUNK + u * 16, raw);
          ptx::tmem_ld_wait();
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
            // row r of a K-major 128B-swizzled tile: 16-byte chunk ch at (ch ^ (r & 7)); head hl = chunks 2 hl, 2 hl + 1
            const uint32_t rowa = (sect == 0 ? sQ : sK) + (uint32_t)r * 128u;
#pragma unroll
            for (int q = 0; q < 2; ++q)
              qa_sts128(rowa + (uint32_t)(((2 * hl + q) ^ (r & 7)) << 4), qa_pack(v[8 * q], v[8 * q + 1]),
                        qa_pack(v[8 * q + 2], v[8 * q + 3]), qa_pack(v[8 * q + 4], v[8 * q + 5]),
                        qa_pack(v[8 * q + 6], v[8 * q + 7]));
          } else {
            // V^T of head hl: element (dim d, key r) of a [16 x 128 B] K-major tile per 64 keys
            const uint32_t keya = sV + (uint32_t)hl * 4096u + (uint32_t)(r >> 6) * 2048u + (uint32_t)(r & 7) * 2u;
            const uint32_t kch = (uint32_t)(r & 63) >> 3;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
              const __nv_bfloat16 hv = __float2bfloat16(v[d]);
              qa_sts16(keya + (uint32_t)d * 128u + ((kch ^ (uint32_t)(d & 7)) << 4), *reinterpret_cast<const uint16_t*>(&hv));
            }
          }
        }
        ptx::fence_proxy_async();                          // generic-proxy stores -> UMMA (async proxy)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(qkv_ready);
        if (PROF) lacc[10] += clock64() - t_drain2;
        // ---- the four heads of the group, one after the other; S / P buffers alternate ----
        uint32_t bsum[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) bsum[j] = 0u;
#pragma unroll 1
        for (int hl = 0; hl < 4; ++hl, ++hcount) {
          const int h = grp * 4 + hl;
          const uint32_t buf = hl & 1;
          const uint32_t t_s = lane_base + T_S + buf * 128u;
          { QA_T0(); ptx::mbar_wait(s_full + 8 * buf, (hcount >> 1) & 1); QA_ACC(11); }
          ptx::tc_fence_after();
          const long long t_sm = PROF ? clock64() : 0;
          uint32_t raw[16];
          ptx::tmem_ld16(t_s + ws * 64 + k0, raw);
          ptx::tmem_ld_wait();
          // shift for the exponent: an upper bound of the row maximum that needs no bias look-ups:
          // max_j(raw) * sc + (largest table sum of the head)
          float mx = __uint_as_float(raw[0]);
#pragma unroll
          for (int j = 1; j < KMAX; ++j) if (j < nk) mx = fmaxf(mx, __uint_as_float(raw[j]));
          float* xch = s_mx + (hcount & 1) * 512;
          xch[part * 128 + r] = mx;
          // every part has its scores in registers behind this barrier: P may now overwrite S
          asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory");
          mx = fmaxf(fmaxf(xch[r], xch[128 + r]), fmaxf(xch[256 + r], xch[384 + r]));
          const float shift = fmaf(mx, sc, s_bmax[h]);
          if ((hl & 1) == 0) {
            // bias of every pair for this head AND the next one (fp16 pair per look-up), kept in registers
            const int hp = h >> 1, Hh = H >> 1;
            const uint32_t tx = tab_u + (uint32_t)(hp * SUBP) * 4u;
            const uint32_t ty = tab_u + (uint32_t)((Hh + hp) * SUBP) * 4u;
            const uint32_t tz = tab_u + (uint32_t)((2 * Hh + hp) * SUBP) * 4u;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
              if (j < nk) {
                const uint32_t cd = code[j];
                bsum[j] = qa_hadd2(qa_hadd2(qa_lds_u32(tx + (cd & 0x3ffu)), qa_lds_u32(ty + ((cd >> 10) & 0x3ffu))),
                                   qa_lds_u32(tz + (cd >> 20)));
              }
            }
          }
          float l = 0.f;
          uint32_t pk[(KMAX + 1) / 2];
#pragma unroll
          for (int j = 0; j < KMAX; j += 2) {
            float pe[2] = {0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (j + e < KMAX && j + e < nk) {
                const float b = (hl & 1) ? qa_half_hi(bsum[j + e]) : qa_half_lo(bsum[j + e]);
                pe[e] = qa_ex2(fmaf(__uint_as_float(raw[j + e]), sc, b - shift));
              }
            }
            l += pe[0] + pe[1];
            pk[j >> 1] = valid ? qa_pack(pe[0], pe[1]) : 0u;
          }
          s_l[(hl * 4 + part) * 128 + r] = l;
          // P (bf16 pairs along the keys) over the S buffer: this part's columns of the own window slot
          // (the last part also clears the tail up to 64 keys) + a quarter of the other slot's zeros
          {
            const uint32_t pc = t_s + ws * 32 + part * (KP / 2);
            if (part < 3) qa_st_cols<KP / 2>(pc, pk);
            else qa_st_cols_tail<(KMAX + 1) / 2, 32 - 3 * (KP / 2)>(pc, pk);
            uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            ptx::tmem_st8(t_s + (ws ^ 1) * 32 + part * 8, z);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(p_ready + 8 * buf);
          if (PROF) lacc[12] += clock64() - t_sm;
        }
        drain_o(gcount, grp, row, valid);
      }
    }
    if (PROF) lacc[15] = clock64() - t_role;
  }
  if (PROF && blockIdx.x == 0) {
    if (threadIdx.x == 16 * 32) for (int i = 0; i < 6; ++i) p.prof[i] = lacc[i];
    if (threadIdx.x == 0) for (int i = 6; i < 16; ++i) p.prof[i] = lacc[i];
    if (threadIdx.x == 12 * 32) for (int i = 6; i < 16; ++i) p.prof[16 + i] = lacc[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
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
    for (int i = 6; i < 16; ++i) fprintf(stderr, " p0:%s=%.1fk", nm[i], h[i] / 1e3);
    for (int i = 6; i < 16; ++i) fprintf(stderr, " p3:%s=%.1fk", nm[i], h[16 + i] / 1e3);
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

int hfl_qkv_attn(const void* y, const void* Wg, const float* bias_g, void* out, const int16_t* xyzb,
                 const float* rpe, int64_t n_win, int64_t rows, int32_t H, int32_t C, int32_t K, int32_t dil,
                 int32_t hat, int32_t bnd, float scale, void* stream_) {
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
