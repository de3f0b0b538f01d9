// Attention cores: head_dim is 16 for every
// stage of every shipped configuration, so one score tile is a 16x8x16 MMA and the
// softmax / relative-position bias dominate -- these kernels keep everything of one
// (window, head) in registers + a warp-private smem slice and never materialise the
// (N_win,K,K) masks or the (N_win,K,K,3) rel-pos tensors of the reference
// (models/octree.py:193-209, 272-283): the batch mask and the RPE index are computed
// from a per-token (x,y,z,submap) int16x4 table.
//   k_window_attn3: stand-alone octree window attention on mma.sync for the window shapes the fused
//                   tcgen05 kernel (attn_fused.cu) does not take (K + relay token > 64: the K = 64
//                   hierarchical levels of the CS-Wild-Places / Campus3D cfgs, K = 96 sweeps), and the
//                   A/B reference of the fused path (HFL_FUSED_ATTN=0): pair bias summed once per 8 heads
//                   into a fragment-ordered smem table
//   k_varlen_attn : relay-token self-attention over ragged per-submap sequences
// qkv is the bf16 output of the tcgen05 projection GEMM, laid out [row, 3C] as
// [q | k | v] x [head, 16]  (octformer_backbone.py:71-72).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int AT_HD = 16;
constexpr int AT_NT = 13;              // score n-tiles (8 keys each) -> up to 104 keys per pass (K = 96 + relay token)
constexpr int AT_KEYS = AT_NT * 8;
constexpr int AT_RS = 48;              // smem row stride in bytes (16 bf16 + pad, conflict-free): ragged kernel
constexpr int AT_ROW = 32;             // window kernel: unpadded 16 x bf16 rows, halves swizzled
constexpr float LOG2E = 1.4426950408889634f;

struct WinAttnParams {
  const __nv_bfloat16* qkv;   // [rows, 3C]
  __nv_bfloat16* out;         // [rows, C]
  const short4* xyzb;         // [n_pad] token table (x,y,z,submap)
  const float* rpe;           // [3*(2*bnd+1), H] or NULL
  int n_win, H, C, K, dil, hat, bnd;
  float scale;
};

// row of slot s of window w, and the token index behind it (-1 for the relay token)
__device__ __forceinline__ void slot_row(const WinAttnParams& p, int w, int s, int64_t& row,
                                         int64_t& tok) {
  if (p.hat) {
    row = (int64_t)w * (p.K + 1) + s;
    tok = s == 0 ? -1 : (int64_t)w * p.K + (s - 1);
  } else if (p.dil > 1) {
    row = (int64_t)(w / p.dil) * p.K * p.dil + (int64_t)s * p.dil + (w % p.dil);
    tok = row;
  } else {
    row = (int64_t)w * p.K + s;
    tok = row;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- fp16 pair helpers ----
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float h_lo(uint32_t v) {
  float f;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.f32.f16 %0, lo;\n\t}" : "=f"(f) : "r"(v));
  return f;
}
__device__ __forceinline__ float h_hi(uint32_t v) {
  float f;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.f32.f16 %0, hi;\n\t}" : "=f"(f) : "r"(v));
  return f;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---------------------------------------------------------------------------
// v3: the bias of one (query, key) pair is summed ONCE for eight heads and parked in shared memory.
// A per-head design spends ~20 thread-instructions per score, half of them on the bias: three field
// extractions + three table look-ups per pair and pair of heads, repeated by every warp.  Here a
// build phase walks the pairs of the window once per group of 8 heads: the RPE tables hold one
// 16-byte entry (8 x fp16, one per head) per offset, so 3 LDS.128 + 8 HADD2 give the bias of a pair for
// all 8 heads; the sums are written in MMA-fragment order ([head][m-tile][n-tile][lane] -> 4 fp16 =
// this lane's four scores), so the score loop of a warp (= one head) costs ONE conflict-free LDS.64
// + 4 converts per score tile.  Masked pairs (other submap / padding key) get -inf for all heads,
// pairs without RPE (relay-token key, disable_RPE) the zero slot.
// H = 16 runs as two passes of 8 heads (K/V of 8 heads staged per pass): ~80 KB of shared memory per
// window at K = 48, two windows per SM.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 hadd2_x4(const uint4 a, const uint4 b) {
  return make_uint4(hadd2_u32(a.x, b.x), hadd2_u32(a.y, b.y), hadd2_u32(a.z, b.z), hadd2_u32(a.w, b.w));
}

struct Win3Smem {            // byte offsets into dynamic shared memory
  int rpe, tok, rtm, prob, bias, kv, total;
};
// compact: only the current pass's RPE table is resident and the relay-token probabilities reuse the
// warp's own (finished) bias slice -- for windows whose full layout exceeds the shared memory (K = 96 + relay token)
__host__ __device__ inline Win3Smem win3_layout(int H, int K, int hat, int bnd, bool compact = false) {
  const int L = K + hat, NT = (L + 7) / 8, NTC = NT * 8, sub = 2 * bnd + 3, passes = H / 8;
  Win3Smem m;
  int o = 0;
  m.rpe = o;  o += (compact ? 1 : passes) * 3 * sub * 16;
  m.tok = o;  o += NTC * 8;
  m.rtm = o;  o += (NTC + 15) & ~15;
  m.prob = o; o += compact ? 0 : 8 * NTC * 4;
  o = (o + 15) & ~15;
  m.bias = o; o += 8 * (K / 16) * NT * 256;
  m.kv = o;   o += 8 * 2 * NTC * AT_ROW;
  m.total = o;
  return m;
}

// WPH = warps per head: 1 -> 8 warps, two windows per SM; 2 -> 16 warps sharing a head's query tiles,
// for windows whose tables allow only one CTA per SM (K = 64: 123 KB)
// (a three-windows-per-SM instantiation, 80 registers + compact tables, was measured in round 2: 28.0 vs
// 25.9 ms per step -- slower, removed)
template <int NT, int WPH, bool COMPACT = false>
__global__ void __launch_bounds__(256 * WPH, WPH == 1 ? 2 : 1) k_window_attn3(const WinAttnParams p) {
  constexpr int NTC = NT * 8;
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp_id = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int warp = warp_id & 7, part = warp_id >> 3;         // head within the pass, share of its query tiles
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K, hat = p.hat, L = K + hat;
  const int num = 2 * p.bnd + 1, sub = num + 2;
  const int passes = p.H >> 3;
  const Win3Smem lay = win3_layout(p.H, K, hat, p.bnd, COMPACT);
  uint4* s_rpe = reinterpret_cast<uint4*>(smem + lay.rpe);                // [passes][3][sub] 8 x fp16
  short4* s_tok = reinterpret_cast<short4*>(smem + lay.tok);              // [NTC]
  uint8_t* s_rtm = smem + lay.rtm;                                        // [NTC] relay-token row mask
  float* s_prob = reinterpret_cast<float*>(smem + lay.prob);              // [8][NTC]
  uint2* s_bias = reinterpret_cast<uint2*>(smem + lay.bias);              // [8][n_mt][NT][32]
  const uint32_t smem_u = ptx::smem_u32(smem);
  const uint32_t kv_u = smem_u + (uint32_t)lay.kv;                        // [8][K | V][NTC x 32 B]
  auto kv_off = [](int row, int half) -> uint32_t {
    return (uint32_t)row * AT_ROW + (uint32_t)((half ^ ((row >> 2) & 1)) << 4);
  };
  const int n_mt = K / 16;

  // RPE tables as fp16 of bias / scale, 8 heads per entry; slot num = 0, slot num + 1 of axis 0 = -inf.
  // The summed bias seeds the accumulator of the QK^T MMA (acc = q.k + bias / scale); the softmax scale
  // is applied inside the exponent: p = 2^(sc * acc - sc * max), one FFMA per score.
  const float inv_scale = 1.0f / p.scale;
  if constexpr (!COMPACT) {
    for (int i = threadIdx.x; i < passes * 3 * sub * 4; i += blockDim.x) {
      const int hp = i & 3, e = i >> 2;               // head pair within the entry, entry index
      const int k = e % sub, axis = (e / sub) % 3, ps = e / (3 * sub);
      float a = 0.f, b = 0.f;
      if (k < num) {
        if (p.rpe) {
          const float* src = p.rpe + (size_t)(axis * num + k) * p.H + ps * 8 + 2 * hp;
          a = __ldg(src) * inv_scale;
          b = __ldg(src + 1) * inv_scale;
        }
      } else if (k == num + 1 && axis == 0) {
        a = b = -INFINITY;
      }
      reinterpret_cast<uint32_t*>(s_rpe)[i] = pack_h2(a, b);
    }
  }
  // compact layout: the table of pass ps only, refilled at the start of every pass
  auto fill_rpe_pass = [&](int ps) {
    for (int i = threadIdx.x; i < 3 * sub * 4; i += blockDim.x) {
      const int hp = i & 3, e = i >> 2;
      const int k = e % sub, axis = e / sub;
      float a = 0.f, b = 0.f;
      if (k < num) {
        if (p.rpe) {
          const float* src = p.rpe + (size_t)(axis * num + k) * p.H + ps * 8 + 2 * hp;
          a = __ldg(src) * inv_scale;
          b = __ldg(src + 1) * inv_scale;
        }
      } else if (k == num + 1 && axis == 0) {
        a = b = -INFINITY;
      }
      reinterpret_cast<uint32_t*>(s_rpe)[i] = pack_h2(a, b);
    }
  };
  for (int i = threadIdx.x; i < NTC; i += blockDim.x)
    if (i >= L) s_tok[i] = make_short4(0, 0, 0, -2);
  const int C3 = 3 * p.C;
  const float sc = p.scale * LOG2E;
  const uint32_t ones[2] = {0x3F803F80u, 0x3F803F80u};
  const int o_zero = num * 16, o_inf = (num + 1) * 16;
  const bool use_rpe = p.rpe != nullptr;

  for (int w = blockIdx.x; w < p.n_win; w += gridDim.x) {
    // row(slot) = row_base + slot * row_step; token(slot) = tok_base + slot (hat: slot - 1, relay -> first token)
    int64_t row_base, tok_base;
    int row_step = 1;
    if (hat) { row_base = (int64_t)w * (K + 1); tok_base = (int64_t)w * K - 1; }
    else if (p.dil > 1) {
      row_base = (int64_t)(w / p.dil) * K * p.dil + (w % p.dil);
      row_step = p.dil;
      tok_base = 0;
    } else { row_base = (int64_t)w * K; tok_base = row_base; }
    for (int ps = 0; ps < passes; ++ps) {
      __syncthreads();                              // previous pass / window fully consumed
      if (ps == 0) {
        for (int sl = threadIdx.x; sl < L; sl += blockDim.x) {
          int64_t tok = (p.dil > 1 && !hat) ? row_base + (int64_t)sl * row_step : tok_base + sl;
          if (hat && sl == 0) tok = (int64_t)w * K;
          ptx::cp_async8(ptx::smem_u32(s_tok + sl), p.xyzb + tok);
        }
      }
      ptx::cp_async_commit();
      // K / V of this pass's 8 heads: 16 pieces of 16 B per (slot, K | V); a half-warp per row,
      // 16 rows per sweep of the CTA
      {
        const int piece = threadIdx.x & 15;
        const __nv_bfloat16* src0 = p.qkv + p.C + ps * 128 + piece * 8;
        const uint32_t dst0 = kv_u + (uint32_t)(piece >> 1) * (2 * NTC * AT_ROW);
        for (int r = threadIdx.x >> 4; r < 2 * NTC; r += 16 * WPH) {
          const int which = r >= NTC, sl = r - which * NTC;            // 0 = K, 1 = V
          const bool ok = sl < L;
          const int64_t row = ok ? row_base + (int64_t)sl * row_step : 0;
          ptx::cp_async16(dst0 + (uint32_t)which * (NTC * AT_ROW) + kv_off(sl, piece & 1),
                          src0 + row * C3 + which * p.C, ok ? 16u : 0u);
        }
      }
      ptx::cp_async_commit();
      // this warp's head; the Q fragments of its first query tile (and the relay-token query row) are
      // requested now, so that their latency hides behind the bias build
      const int h = ps * 8 + warp;
      const __nv_bfloat16* qh = p.qkv + h * AT_HD;
      // Q / output rows of this warp's query tiles advance by a constant stride: running pointers
      // (the 64-bit row arithmetic per tile was ~10 % of the head loop's instructions)
      const int64_t qrow0 = row_base + (int64_t)(part * 16 + g + hat) * row_step;
      const uint32_t* qp0 = reinterpret_cast<const uint32_t*>(qh + qrow0 * C3) + t;
      const size_t q_half = (size_t)4 * row_step * C3, q_tile = (size_t)8 * WPH * row_step * C3;   // in 32-bit words
      uint32_t* op0 = reinterpret_cast<uint32_t*>(p.out + qrow0 * p.C + h * AT_HD) + t;
      const size_t o_half = (size_t)4 * row_step * p.C, o_tile = (size_t)8 * WPH * row_step * p.C;
      auto load_q = [&](uint32_t (&qf)[4]) {
        qf[0] = __ldg(qp0); qf[2] = __ldg(qp0 + 4);
        qf[1] = __ldg(qp0 + q_half); qf[3] = __ldg(qp0 + q_half + 4);
        qp0 += q_tile;
      };
      uint32_t qn[4] = {0u, 0u, 0u, 0u};
      if (part < n_mt) load_q(qn);
      uint4 qrt0 = make_uint4(0u, 0u, 0u, 0u), qrt1 = qrt0;
      if (hat && part == 0) {
        const uint4* qp = reinterpret_cast<const uint4*>(qh + row_base * C3);
        qrt0 = __ldg(qp);
        qrt1 = __ldg(qp + 1);
      }
      if constexpr (COMPACT) fill_rpe_pass(ps);     // this pass's table only (previous pass fully consumed)
      if (ps == 0) {
        ptx::cp_async_wait<1>();                    // tokens landed (K / V still in flight)
        __syncthreads();
        for (int j = threadIdx.x; j < NTC; j += blockDim.x) s_rtm[j] = s_tok[0].w == s_tok[j].w;
      } else if constexpr (COMPACT) {
        __syncthreads();                            // the refilled table is visible to the build
      }
      // ---- bias of every (query, key) pair for the 8 heads of this pass, in fragment order ----
      {
        const uint8_t* tab = reinterpret_cast<const uint8_t*>(s_rpe + (COMPACT ? 0 : ps) * 3 * sub);
        for (int tile = warp_id; tile < n_mt * NT; tile += 8 * WPH) {
          const int mt = tile / NT, nt = tile - mt * NT;
          const short4 ti[2] = {s_tok[mt * 16 + g + hat], s_tok[mt * 16 + g + 8 + hat]};
          const int c0 = nt * 8 + 2 * t;
          const short4 tj[2] = {s_tok[c0], s_tok[c0 + 1]};
          uint4 sum[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const short4 a = ti[e >> 1], b = tj[e & 1];
            int ox = o_inf, oy = o_zero, oz = o_zero;
            if (a.w == b.w) {
              ox = o_zero;
              if (use_rpe && !(hat && c0 + (e & 1) == 0)) {
                ox = (min(max((int)a.x - (int)b.x, -p.bnd), p.bnd) + p.bnd) * 16;
                oy = (min(max((int)a.y - (int)b.y, -p.bnd), p.bnd) + p.bnd) * 16;
                oz = (min(max((int)a.z - (int)b.z, -p.bnd), p.bnd) + p.bnd) * 16;
              }
            }
            const uint4 bx = *reinterpret_cast<const uint4*>(tab + ox);
            const uint4 by = *reinterpret_cast<const uint4*>(tab + sub * 16 + oy);
            const uint4 bz = *reinterpret_cast<const uint4*>(tab + 2 * sub * 16 + oz);
            sum[e] = hadd2_x4(hadd2_x4(by, bz), bx);
          }
          uint2* dst = s_bias + (size_t)(mt * NT + nt) * 32 + lane;
          const size_t hstride = (size_t)n_mt * NT * 32;
          const uint32_t* s0 = reinterpret_cast<const uint32_t*>(&sum[0]);
          const uint32_t* s1 = reinterpret_cast<const uint32_t*>(&sum[1]);
          const uint32_t* s2 = reinterpret_cast<const uint32_t*>(&sum[2]);
          const uint32_t* s3 = reinterpret_cast<const uint32_t*>(&sum[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            dst[(2 * k) * hstride] = make_uint2(__byte_perm(s0[k], s1[k], 0x5410), __byte_perm(s2[k], s3[k], 0x5410));
            dst[(2 * k + 1) * hstride] = make_uint2(__byte_perm(s0[k], s1[k], 0x7632), __byte_perm(s2[k], s3[k], 0x7632));
          }
        }
      }
      ptx::cp_async_wait<0>();
      __syncthreads();

      // ---- warp = head: K/16 query tiles ----
      const uint32_t sK_u = kv_u + (uint32_t)warp * (2 * NTC * AT_ROW);
      const uint32_t sV_u = sK_u + NTC * AT_ROW;
      const uint8_t* sK = smem + lay.kv + (size_t)warp * (2 * NTC * AT_ROW);
      const uint2* bias_h = s_bias + (size_t)warp * n_mt * NT * 32 + lane;
      // this head's K fragments are the same for every query tile: keep them in registers
      uint32_t kf[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        kf[nt][0] = *reinterpret_cast<const uint32_t*>(sK + kv_off(nt * 8 + g, 0) + t * 4);
        kf[nt][1] = *reinterpret_cast<const uint32_t*>(sK + kv_off(nt * 8 + g, 1) + t * 4);
      }
      for (int mt = part; mt < n_mt; mt += WPH) {
        uint32_t qa[4] = {qn[0], qn[1], qn[2], qn[3]};
        if (mt + WPH < n_mt) load_q(qn);             // next tile's Q in flight during this tile
        float s[NT][4];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint2 bb = bias_h[(mt * NT + nt) * 32];
          s[nt][0] = h_lo(bb.x); s[nt][1] = h_hi(bb.x);
          s[nt][2] = h_lo(bb.y); s[nt][3] = h_hi(bb.y);
          ptx::mma16816(s[nt], qa, kf[nt]);
          mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        mx0 *= sc;                                   // exponent = sc * acc - sc * max
        mx1 *= sc;
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kt = 0; kt < (NT + 1) / 2; ++kt) {
          uint32_t pa[4];
          pa[0] = pack_bf16(fast_exp2(fmaf(s[2 * kt][0], sc, -mx0)), fast_exp2(fmaf(s[2 * kt][1], sc, -mx0)));
          pa[1] = pack_bf16(fast_exp2(fmaf(s[2 * kt][2], sc, -mx1)), fast_exp2(fmaf(s[2 * kt][3], sc, -mx1)));
          if (2 * kt + 1 < NT) {
            pa[2] = pack_bf16(fast_exp2(fmaf(s[2 * kt + 1][0], sc, -mx0)), fast_exp2(fmaf(s[2 * kt + 1][1], sc, -mx0)));
            pa[3] = pack_bf16(fast_exp2(fmaf(s[2 * kt + 1][2], sc, -mx1)), fast_exp2(fmaf(s[2 * kt + 1][3], sc, -mx1)));
          } else {
            pa[2] = pa[3] = 0u;
          }
          uint32_t vb[4];
          const int mi = lane >> 3;
          int key = kt * 16 + (mi & 1) * 8 + (lane & 7);
          key = key < NTC ? key : 0;
          ptx::ldmatrix_x4_trans(vb, sV_u + kv_off(key, mi >> 1));
          uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
          ptx::mma16816(o[0], pa, b0);
          ptx::mma16816(o[1], pa, b1);
          ptx::mma16816(ls, pa, ones);
        }
        const float i0 = 1.f / ls[0], i1 = 1.f / ls[2];
        op0[0] = pack_bf16(o[0][0] * i0, o[0][1] * i0);
        op0[4] = pack_bf16(o[1][0] * i0, o[1][1] * i0);
        op0[o_half] = pack_bf16(o[0][2] * i1, o[0][3] * i1);
        op0[o_half + 4] = pack_bf16(o[1][2] * i1, o[1][3] * i1);
        op0 += o_tile;
      }
      // ---- the relay-token query row of this head (no RPE): lanes over keys, then over (dim, half) ----
      if (hat && part == 0) {
        const int64_t rowq = row_base;
        const uint8_t* sV = sK + NTC * AT_ROW;
        float q[AT_HD];
        {
          // keep the unpacking of the prefetched row HERE: without the (empty) volatile asm the compiler
          // hoists it to right behind the loads and the warp waits for them before the bias build
          asm volatile("" : "+r"(qrt0.x), "+r"(qrt0.y), "+r"(qrt0.z), "+r"(qrt0.w),
                            "+r"(qrt1.x), "+r"(qrt1.y), "+r"(qrt1.z), "+r"(qrt1.w));
          const uint4 a = qrt0, b = qrt1;
          const uint32_t wds[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[i]));
            q[2 * i] = f.x; q[2 * i + 1] = f.y;
          }
        }
        float sj[(NTC + 31) / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < (NTC + 31) / 32; ++u) {
          const int j = lane + 32 * u;
          float a = -INFINITY;
          if (j < NTC && s_rtm[j]) {
            const uint4 ka = *reinterpret_cast<const uint4*>(sK + kv_off(j, 0));
            const uint4 kb = *reinterpret_cast<const uint4*>(sK + kv_off(j, 1));
            const uint32_t wds[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
            a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[i]));
              a = fmaf(q[2 * i], f.x, a);
              a = fmaf(q[2 * i + 1], f.y, a);
            }
            a *= sc;
          }
          sj[u] = a;
          mx = fmaxf(mx, a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float l = 0.f;
        // compact: this warp's own bias slice is dead once its query tiles are done (WPH == 1)
        float* pr = COMPACT ? reinterpret_cast<float*>(s_bias + (size_t)warp * n_mt * NT * 32) : s_prob + warp * NTC;
#pragma unroll
        for (int u = 0; u < (NTC + 31) / 32; ++u) {
          const int j = lane + 32 * u;
          const float e = fast_exp2(sj[u] - mx);
          l += e;
          if (j < NTC) pr[j] = e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        __syncwarp();
        const int d = lane & 15, half = lane >> 4;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
        const uint8_t* vcol = sV + (d & 7) * 2;
#pragma unroll 2
        for (int j0 = half; j0 < NTC; j0 += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 2 * u;
            acc4[u] = fmaf(pr[j], __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(vcol + kv_off(j, d >> 3))), acc4[u]);
          }
        }
        float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (half == 0) p.out[rowq * p.C + h * AT_HD + d] = __float2bfloat16(acc / l);
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Ragged (per-submap) self-attention for the relay tokens: flash-style loop over
// key blocks of 80; CTA = (submap, head, chunk of 256 query rows), 4 warps x 4 m-tiles.
// ---------------------------------------------------------------------------
struct VarAttnParams {
  const __nv_bfloat16* qkv;   // [total, 3C] compact, submap-major
  __nv_bfloat16* out;         // [total, C]
  const int32_t* cu;          // [B+1] sequence offsets
  const int32_t* ids;         // [total] tokens attend iff ids equal
  int B, H, C;
  float scale;
};
constexpr int VA_NT = 10;       // key tiles per block of the ragged kernel (even: the P.V loop takes pairs)
constexpr int VA_KEYS = VA_NT * 8;
constexpr int VA_WARPS = 4;
constexpr int VA_MT = 4;        // m-tiles per warp -> 256 query rows per CTA

__global__ void __launch_bounds__(VA_WARPS * 32) k_varlen_attn(const VarAttnParams p) {
  __shared__ __align__(16) uint8_t sK[VA_KEYS * AT_RS];
  __shared__ __align__(16) uint8_t sV[VA_KEYS * AT_RS];
  __shared__ int32_t sId[VA_KEYS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x, h = blockIdx.y;
  const int beg = p.cu[b], L = p.cu[b + 1] - beg;
  const int q0 = blockIdx.z * (VA_WARPS * VA_MT * 16);
  if (q0 >= L) return;
  const int C3 = 3 * p.C;
  const float sc = p.scale * LOG2E;
  const uint32_t sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);

  uint32_t qa[VA_MT][4];
  int32_t id0[VA_MT], id1[VA_MT];
  float m0[VA_MT], m1[VA_MT], l0[VA_MT], l1[VA_MT], o[VA_MT][2][4];
#pragma unroll
  for (int u = 0; u < VA_MT; ++u) {
    const int i0 = q0 + (warp * VA_MT + u) * 16 + g, i1 = i0 + 8;
    qa[u][0] = qa[u][1] = qa[u][2] = qa[u][3] = 0u;
    id0[u] = id1[u] = -2;
    if (i0 < L) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p.qkv + (int64_t)(beg + i0) * C3 + h * AT_HD);
      qa[u][0] = __ldg(q + t); qa[u][2] = __ldg(q + t + 4);
      id0[u] = __ldg(p.ids + beg + i0);
    }
    if (i1 < L) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p.qkv + (int64_t)(beg + i1) * C3 + h * AT_HD);
      qa[u][1] = __ldg(q + t); qa[u][3] = __ldg(q + t + 4);
      id1[u] = __ldg(p.ids + beg + i1);
    }
    m0[u] = m1[u] = -INFINITY;
    l0[u] = l1[u] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[u][a][c] = 0.f;
  }

  for (int kb0 = 0; kb0 < L; kb0 += VA_KEYS) {
    __syncthreads();
    for (int s = threadIdx.x; s < VA_KEYS; s += blockDim.x) {
      const bool ok = kb0 + s < L;
      const __nv_bfloat16* src = p.qkv + (int64_t)(beg + (ok ? kb0 + s : 0)) * C3 + p.C + h * AT_HD;
      ptx::cp_async16(sK_u + s * AT_RS, src, ok ? 16u : 0u);
      ptx::cp_async16(sK_u + s * AT_RS + 16, src + 8, ok ? 16u : 0u);
      ptx::cp_async16(sV_u + s * AT_RS, src + p.C, ok ? 16u : 0u);
      ptx::cp_async16(sV_u + s * AT_RS + 16, src + p.C + 8, ok ? 16u : 0u);
      sId[s] = ok ? __ldg(p.ids + beg + kb0 + s) : -1;
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int u = 0; u < VA_MT; ++u) {
      if (q0 + (warp * VA_MT + u) * 16 >= L) continue;        // warp-uniform
      float s[VA_NT][4];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < VA_NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        uint32_t kb[2];
        const uint8_t* kr = sK + (nt * 8 + g) * AT_RS + t * 4;
        kb[0] = *reinterpret_cast<const uint32_t*>(kr);
        kb[1] = *reinterpret_cast<const uint32_t*>(kr + 16);
        ptx::mma16816(s[nt], qa[u], kb);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idj = sId[nt * 8 + 2 * t + e];
          s[nt][e] = (idj == id0[u]) ? s[nt][e] * sc : -INFINITY;
          s[nt][2 + e] = (idj == id1[u]) ? s[nt][2 + e] * sc : -INFINITY;
          mx0 = fmaxf(mx0, s[nt][e]);
          mx1 = fmaxf(mx1, s[nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float n0 = fmaxf(m0[u], mx0), n1 = fmaxf(m1[u], mx1);
      const float e0 = n0 == -INFINITY ? 0.f : n0, e1 = n1 == -INFINITY ? 0.f : n1;
      const float c0 = exp2f(m0[u] - e0), c1 = exp2f(m1[u] - e1);   // exp2(-inf) = 0 on first block
      m0[u] = n0; m1[u] = n1;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < VA_NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - e0); s[nt][1] = exp2f(s[nt][1] - e0);
        s[nt][2] = exp2f(s[nt][2] - e1); s[nt][3] = exp2f(s[nt][3] - e1);
        a0 += s[nt][0] + s[nt][1];
        a1 += s[nt][2] + s[nt][3];
      }
      a0 += __shfl_xor_sync(0xffffffffu, a0, 1); a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
      l0[u] = l0[u] * c0 + a0;
      l1[u] = l1[u] * c1 + a1;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        o[u][a][0] *= c0; o[u][a][1] *= c0; o[u][a][2] *= c1; o[u][a][3] *= c1;
      }
#pragma unroll
      for (int kt = 0; kt < VA_NT / 2; ++kt) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
        pa[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
        pa[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        pa[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
        uint32_t vb[4];
        const int mi = lane >> 3;
        const uint32_t va = sV_u + (kt * 16 + (mi & 1) * 8 + (lane & 7)) * AT_RS + (mi >> 1) * 16;
        ptx::ldmatrix_x4_trans(vb, va);
        uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
        ptx::mma16816(o[u][0], pa, b0);
        ptx::mma16816(o[u][1], pa, b1);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < VA_MT; ++u) {
    const int i0 = q0 + (warp * VA_MT + u) * 16 + g, i1 = i0 + 8;
    const float r0 = l0[u] > 0.f ? 1.f / l0[u] : 0.f, r1 = l1[u] > 0.f ? 1.f / l1[u] : 0.f;
    if (i0 < L) {
      uint32_t* d = reinterpret_cast<uint32_t*>(p.out + (int64_t)(beg + i0) * p.C + h * AT_HD);
      d[t] = pack_bf16(o[u][0][0] * r0, o[u][0][1] * r0);
      d[t + 4] = pack_bf16(o[u][1][0] * r0, o[u][1][1] * r0);
    }
    if (i1 < L) {
      uint32_t* d = reinterpret_cast<uint32_t*>(p.out + (int64_t)(beg + i1) * p.C + h * AT_HD);
      d[t] = pack_bf16(o[u][0][2] * r1, o[u][0][3] * r1);
      d[t + 4] = pack_bf16(o[u][1][2] * r1, o[u][1][3] * r1);
    }
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_window_attn(const void* qkv, void* out, const int16_t* xyzb, const float* rpe,
                    int64_t n_win, int32_t H, int32_t C, int32_t K, int32_t dil, int32_t hat,
                    int32_t bnd, float scale, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_win == 0) return HFL_OK;
  HFL_CHECK_ARG(qkv && out && xyzb, "null argument");
  HFL_CHECK_ARG(C == H * AT_HD && H <= 16, "head_dim must be 16, at most 16 heads");
  HFL_CHECK_ARG(K % 16 == 0 && K + (hat ? 1 : 0) <= AT_KEYS, "window must be a multiple of 16 and (+relay token) fit 104 keys");
  HFL_CHECK_ARG(dil >= 1 && (!hat || dil == 1), "dilation is not used with relay tokens");
  HFL_CHECK_ARG(n_win % dil == 0, "window count must be a multiple of the dilation");
  HFL_CHECK_ARG(bnd >= 0, "bad RPE bound");
  WinAttnParams p;
  p.qkv = (const __nv_bfloat16*)qkv; p.out = (__nv_bfloat16*)out; p.xyzb = (const short4*)xyzb;
  p.rpe = rpe; p.n_win = (int)n_win; p.H = H; p.C = C; p.K = K; p.dil = dil; p.hat = hat;
  p.bnd = bnd; p.scale = scale;
  const int L = K + (hat ? 1 : 0);
  const int NT = (L + 7) / 8;
  HFL_CHECK_ARG(H % 8 == 0, "the window attention kernel processes heads in groups of 8");
  Win3Smem lay3 = win3_layout(H, K, hat, bnd);
  const bool compact = lay3.total > 227 * 1024;
  if (compact) lay3 = win3_layout(H, K, hat, bnd, true);
  HFL_CHECK_ARG(lay3.total <= 227 * 1024, "window attention tables exceed shared memory");
  // one CTA per SM only (tables > half of the shared memory): 16 warps per window instead of 8
  static const char* w2 = getenv("HFL_ATTN_WPH");
  const bool wph2 = compact || (w2 ? w2[0] == '2' : (2 * lay3.total + 2048 > 227 * 1024 && K >= 32));
  const int smem3 = lay3.total;
  const int per_sm = wph2 ? 1 : 2, sms = sm_count();
  const int grid3 = (int)(n_win < (int64_t)per_sm * sms ? n_win : per_sm * sms);
#define HFL_WA3_CASE(NT_)                                                                         \
  case NT_: {                                                                                     \
    if (compact) {                                                                                \
      HFL_ENSURE_SMEM(smem3, k_window_attn3<NT_, 2, true>);                                       \
      HFL_LAUNCH((k_window_attn3<NT_, 2, true><<<grid3, 512, smem3, st>>>(p)));                   \
    } else if (wph2) {                                                                            \
      HFL_ENSURE_SMEM(smem3, k_window_attn3<NT_, 2>);                                             \
      HFL_LAUNCH((k_window_attn3<NT_, 2><<<grid3, 512, smem3, st>>>(p)));                         \
    } else {                                                                                      \
      HFL_ENSURE_SMEM(smem3, k_window_attn3<NT_, 1>);                                             \
      HFL_LAUNCH((k_window_attn3<NT_, 1><<<grid3, 256, smem3, st>>>(p)));                         \
    }                                                                                             \
    return HFL_OK;                                                                                \
  }
  switch (NT) {
    HFL_WA3_CASE(2) HFL_WA3_CASE(3) HFL_WA3_CASE(4) HFL_WA3_CASE(5) HFL_WA3_CASE(6) HFL_WA3_CASE(7)
    HFL_WA3_CASE(8) HFL_WA3_CASE(9) HFL_WA3_CASE(10) HFL_WA3_CASE(11) HFL_WA3_CASE(12) HFL_WA3_CASE(13)
    default: return fail(HFL_ERR_UNSUPPORTED, "unsupported window size%s (%lld)", "", (long long)K);
  }
#undef HFL_WA3_CASE
}

int hfl_varlen_attn(const void* qkv, void* out, const int32_t* cu_seqlens, const int32_t* ids,
                    int32_t B, int32_t max_len, int32_t H, int32_t C, float scale, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B == 0 || max_len == 0) return HFL_OK;
  HFL_CHECK_ARG(qkv && out && cu_seqlens && ids, "null argument");
  HFL_CHECK_ARG(C == H * AT_HD, "head_dim must be 16");
  VarAttnParams p;
  p.qkv = (const __nv_bfloat16*)qkv; p.out = (__nv_bfloat16*)out; p.cu = cu_seqlens; p.ids = ids;
  p.B = B; p.H = H; p.C = C; p.scale = scale;
  dim3 grid(B, H, (unsigned)ceil_div(max_len, VA_WARPS * VA_MT * 16));
  HFL_LAUNCH((k_varlen_attn<<<grid, VA_WARPS * 32, 0, st>>>(p)));
  return HFL_OK;
}

}  // extern "C"
