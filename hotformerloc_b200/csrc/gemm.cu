// Gather-GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), the workhorse of the
// dense part of the hot path:
//
//   out[m, :] = epilogue( sum_{kk < KD} A[idx[m, kk], :] . W[:, kk*Cin : (kk+1)*Cin]^T )
//
// * KD == 1, idx == NULL  : plain token-major Linear (qkv / proj / fc1 / fc2 / mixer)
// * KD == 27 or 8         : ocnn OctreeConv as an implicit GEMM -- the im2col buffer of
//                           the reference (rows x kdim x Cin) is never materialised;
//                           rows are gathered straight into the swizzled smem operand.
// Persistent, warp-specialised CTA (416 threads, 1 CTA / SM):
//   warps 0-7  epilogue   (TMEM lane quadrant = warp % 4, column half = warp / 4;
//              thread == output row)
//   warp  8    TMEM alloc + single-thread tcgen05.mma issue (M=128, N=block_n, K=16)
//   warps 9-12 A producers: thread == tile row, 8 x 16 B cp.async (zero-fill for
//              neigh < 0 / tail rows) into the 128B-swizzled K-major layout;
//              first producer thread also issues the TMA load of the weight tile.
// 4-stage smem ring (A 16 KB + B <= 32 KB per stage), double-buffered TMEM accumulators
// (2 x block_n columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue fusions: +bias, +residual (fp32 stream), GELU, LayerNorm over the full row
// (thread-local: one thread owns one row of TMEM), ReLU, fp32 and/or bf16 stores with an
// optional output row map (relay-token rows / hierarchical window layout).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_A_BYTES = G_BM * G_BK * 2;    // 16 KB
constexpr int G_B_BYTES = 256 * G_BK * 2;     // 32 KB (block_n <= 256)
constexpr int G_THREADS = 448;       // 8 epilogue + 1 MMA + 4 producer warps + 1 weight-TMA warp
constexpr int G_PROD_THREADS = 128;
constexpr int G_W_MMA = 8, G_W_PROD = 9, G_W_TMA = 13;
constexpr int G_TMA_ISSUERS = 2;     // a TMA load costs its issuing thread ~340 ns (tools/micro/tma_issue.cu)
// streaming mode: 4 stages of (A 16 KB + B 32 KB); weight-stationary mode (Ktot*block_n*2 <=
// 128 KB): the whole W tile lives in smem for the lifetime of the CTA and 4 stages of A stream.
constexpr int G_STAGES_STREAM = 4;
constexpr int G_STAGES_WS = 4;
constexpr int G_WS_W_BYTES = 128 * 1024;
constexpr int G_PIPE_BYTES = 192 * 1024;      // = 4*(16+32) KB (stream) = 4*16 KB + 128 KB (WS)
constexpr int G_STAGE_BYTES = 8 * 2048;       // per epilogue warp: 32 rows x 64 B transpose buffer
constexpr int G_VEC_BYTES = 2 * 3 * 128 * 4;   // per column half: bias | gamma | beta
constexpr int G_SMEM = G_PIPE_BYTES + G_STAGE_BYTES + 256 + 2048 + G_VEC_BYTES;   // + barriers + LN exchange

struct GemmParams {
  const __nv_bfloat16* A;   // [rows_A, Cin]
  const int32_t* idx;       // [M, KD] row gather table or NULL (identity, KD == 1)
  int M, N, KD, Cin;        // Ktot = KD * Cin
  int block_n, n_tiles;
  int ws;                   // weight-stationary mode
  int dense;                // idx == NULL, KD == 1: A tiles are plain 2-D boxes, loaded by TMA (tmap_a)
  int plain2;               // two-chunk fast path for the +bias / bf16-store epilogue (HFL_GEMM_PLAIN2=0 disables)
  // epilogue
  const float* bias;        // [N] or NULL
  const float* res;         // fp32 residual, row-mapped like out_v, or NULL
  int act;                  // 0 none, 1 GELU(erf) on v
  float* out_v_f32;         // v = acc + bias + res     (row-mapped)
  __nv_bfloat16* out_v_bf16;
  const float* ln_g;        // LayerNorm over the row (needs block_n == N) or NULL
  const float* ln_b;
  int relu;                 // ReLU after LayerNorm
  int y_mapped;             // 1: y rows use out_rows, 0: y rows are GEMM rows
  float* out_y_f32;
  __nv_bfloat16* out_y_bf16;
  const int32_t* out_rows;  // [M] output row map or NULL; negative = skip row
  float ln_eps;
};

// Exact (erf) GELU, written for an epilogue that is bound by the FP32 pipe:
//   GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) = h + t - t * erfc(sqrt(2) t),   h = x / 2, t = |h|
// with erfc(sqrt(2) t) = 2^(t P(t)), P a degree-4 minimax fit (8 FP32 ops + one MUFU.EX2 per
// element, no division).  |GELU error| <= 9e-7 over all x evaluated in fp32 (checked against
// scipy.special.erf, tools/fit_gelu.py) -- < 1 % of one bf16 ulp of the stored activation.
__device__ __forceinline__ float gelu_erf(float x) {
  const float h = 0.5f * x, t = fabsf(h);
  float q = fmaf(-1.5639575e-2f, t, 1.1524727e-1f);
  q = fmaf(q, t, -4.1725174e-1f);
  q = fmaf(q, t, -1.8383462f);
  q = fmaf(q, t, -2.3020070f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * t));
  return fmaf(-t, e, h + t);
}

// ---- coalesced epilogue I/O --------------------------------------------------------
// TMEM hands every thread one output ROW; a naive store makes each warp instruction touch
// 32 rows x 16 B (half-empty sectors, 2x the L2 write transactions).  Each epilogue warp
// therefore owns a 32-row x 64-byte transpose buffer (XOR-swizzled, conflict-free both
// ways): threads deposit their row segment, then lanes (row = lane/4 + 8i, seg = lane%4)
// move full 64-byte row segments to / from global memory.
__device__ __forceinline__ uint32_t stage_addr(uint32_t stage, int row, int seg) {
  return stage + (uint32_t)row * 64u + (uint32_t)((seg ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
// one 64-byte-per-row unit: this thread's 16 words w[0..15] -> global rows orow (per lane)
__device__ __forceinline__ void store_unit(uint32_t stage, int lane, const uint32_t* w,
                                           char* gbase, int32_t orow, size_t row_bytes,
                                           size_t col_byte) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    sts128(stage_addr(stage, lane, q), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = (lane >> 2) + 8 * i, seg = lane & 3;
    const uint4 v = lds128(stage_addr(stage, row, seg));
    const int32_t orr = __shfl_sync(0xffffffffu, orow, row);
    if (orr >= 0)
      *reinterpret_cast<uint4*>(gbase + (size_t)orr * row_bytes + col_byte + seg * 16) = v;
  }
}
__device__ __forceinline__ void store_f32x32(uint32_t stage, int lane, float* base, int32_t orow,
                                             int ld, int col, const float (&v)[32]) {
  uint32_t w[16];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(v[u * 16 + j]);
    store_unit(stage, lane, w, reinterpret_cast<char*>(base), orow, (size_t)ld * 4,
               (size_t)(col + u * 16) * 4);
  }
}
__device__ __forceinline__ void store_bf16x32(uint32_t stage, int lane, __nv_bfloat16* base,
                                              int32_t orow, int ld, int col, const float (&v)[32]) {
  uint32_t w[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  store_unit(stage, lane, w, reinterpret_cast<char*>(base), orow, (size_t)ld * 2, (size_t)col * 2);
}
// coalesced fetch of a 32-row x 32-column fp32 chunk: issue (global -> regs) ...
__device__ __forceinline__ void res_issue(const float* base, int32_t orow, int ld, int col, int lane,
                                          uint4 (&buf)[8]) {
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (lane >> 2) + 8 * i, seg = lane & 3;
      const int32_t orr = __shfl_sync(0xffffffffu, orow, row);
      buf[u * 4 + i] = orr >= 0 ? *reinterpret_cast<const uint4*>(base + (size_t)orr * ld + col + u * 16 + seg * 4)
                                : make_uint4(0u, 0u, 0u, 0u);
    }
}
// ... then transpose through the staging buffer and add to this thread's row
__device__ __forceinline__ void res_add(uint32_t stage, int lane, const uint4 (&buf)[8], float (&v)[32]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (lane >> 2) + 8 * i, seg = lane & 3;
      const uint4 b = buf[u * 4 + i];
      sts128(stage_addr(stage, row, seg), b.x, b.y, b.z, b.w);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 b = lds128(stage_addr(stage, lane, q));
      v[u * 16 + 4 * q] += __uint_as_float(b.x);
      v[u * 16 + 4 * q + 1] += __uint_as_float(b.y);
      v[u * 16 + 4 * q + 2] += __uint_as_float(b.z);
      v[u * 16 + 4 * q + 3] += __uint_as_float(b.w);
    }
  }
}

// D320 (HFL_GEMM_DENSE320=1, dense GEMMs only): the same kernel without the four cp.async producer warps --
// 320 threads (8 epilogue + MMA + TMA), which lifts the per-thread register cap from 128 to 168 and lets the
// residual / LayerNorm epilogue keep the next accumulator chunk and TWO residual chunks in flight.
// Measured in round 2 (A/B on one B200): gather_gemm family 35.2 -> 34.0 ms per step; since the residual
// epilogues moved into the fused proj + MLP kernel the difference is within noise -- kept as a toggle.
constexpr int G_THREADS_D320 = 320;
template <bool WS, bool D320 = false>
__global__ void __launch_bounds__(D320 ? G_THREADS_D320 : G_THREADS, 1)
k_gather_gemm(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
              const GemmParams p) {
  constexpr int G_STAGES = WS ? G_STAGES_WS : G_STAGES_STREAM;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if (base & 1023u) __trap();                           // SWIZZLE_128B operands need 1024 B alignment
  uint8_t* smem = smem_raw;
  const uint32_t sA = base;
  const uint32_t sB = base + G_STAGES * G_A_BYTES;      // stream: B ring; WS: resident W tile
  const uint32_t sStage = base + G_PIPE_BYTES;          // epilogue transpose buffers
  const uint32_t sBar = sStage + G_STAGE_BYTES;
  const uint32_t bar_full = sBar;                       // G_STAGES x 8 B
  const uint32_t bar_empty = sBar + 8 * G_STAGES;       // G_STAGES x 8 B
  const uint32_t bar_tfull = sBar + 16 * G_STAGES;      // 2 x 8 B
  const uint32_t bar_tempty = bar_tfull + 16;           // 2 x 8 B
  const uint32_t bar_w = bar_tempty + 16;               // 8 B (WS: W tile landed)
  const uint32_t s_tmem = bar_w + 8;                    // 4 B
  float2* s_ln = reinterpret_cast<float2*>(smem + (sBar + 256 - base));   // [2 halves][128 rows]
  float* s_vec = reinterpret_cast<float*>(smem + (sBar + 256 + 2048 - base));   // [2 halves][3][128]
  volatile uint32_t* tmem_ptr_s =
      reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = D320 ? 9 : G_W_TMA;             // D320: warps 0-7 epilogue, 8 MMA, 9 TMA
  const int m_tiles = (p.M + G_BM - 1) / G_BM;
  const int k_blocks = (p.KD * p.Cin) / G_BK;
  const int BN = p.block_n;
  // tile schedule.  stream: t = m_blk * n_tiles + n_blk, CTAs stride over t.
  // WS: the CTA owns one n_blk (its W tile) and strides over the M tiles.
  const int t_begin = WS ? (int)(blockIdx.x / p.n_tiles) : (int)blockIdx.x;
  const int t_step = WS ? (int)(gridDim.x / p.n_tiles) : (int)gridDim.x;
  const int t_end = WS ? m_tiles : m_tiles * p.n_tiles;
  const int ws_n_blk = blockIdx.x % p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      // arrivals per stage: the producers' cp.async groups (+ the weight TMA's expect-tx in stream
      // mode); dense mode: only the TMA issuer's expect-tx (A box, and the W box in stream mode)
      ptx::mbar_init(bar_full + 8 * s, p.dense ? 1 : (WS ? G_PROD_THREADS : G_PROD_THREADS + 1));
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_w, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == G_W_MMA) {
    ptx::tmem_alloc(s_tmem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == W_TMA) {
    // ===================== weight TMA issuer =====================
    // Kept off the cp.async producers: the ~340 ns a thread spends per TMA instruction would sit
    // on the critical path of the A gather (stream-mode convs run at ~0.5 us per K block).
    if (lane < G_TMA_ISSUERS) {
      ptx::prefetch_tmap(&tmap_w);
      if (p.dense) ptx::prefetch_tmap(&tmap_a);
      const uint32_t b_bytes = (uint32_t)BN * G_BK * 2;
      if (WS) {
        if (lane == 0) {                                   // resident W tile: one shot
          ptx::mbar_arrive_expect_tx(bar_w, (uint32_t)k_blocks * b_bytes);
          for (int kb = 0; kb < k_blocks; ++kb)
            ptx::tma_load_2d(sB + kb * b_bytes, &tmap_w, bar_w, kb * G_BK, ws_n_blk * BN);
        }
        if (p.dense) {
          // dense Linear: the 128 x 64 A box of every K block comes by TMA as well (rows >= M are
          // zero-filled by the tensor map); the cp.async producer warps stay idle
          uint32_t g = 0;
          for (int t = t_begin; t < t_end; t += t_step) {
            for (int kb = 0; kb < k_blocks; ++kb, ++g) {
              if ((int)(g % G_TMA_ISSUERS) != lane) continue;
              const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
              ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
              ptx::mbar_arrive_expect_tx(bar_full + 8 * s, G_A_BYTES);
              ptx::tma_load_2d(sA + s * G_A_BYTES, &tmap_a, bar_full + 8 * s, kb * G_BK, t * G_BM);
            }
          }
        }
      } else {
        uint32_t g = 0;
        for (int t = t_begin; t < t_end; t += t_step) {
          const int n_blk = t % p.n_tiles, m_blk = t / p.n_tiles;
          for (int kb = 0; kb < k_blocks; ++kb, ++g) {
            if ((int)(g % G_TMA_ISSUERS) != lane) continue;
            const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
            ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
            ptx::mbar_arrive_expect_tx(bar_full + 8 * s, p.dense ? b_bytes + G_A_BYTES : b_bytes);
            ptx::tma_load_2d(sB + s * G_B_BYTES, &tmap_w, bar_full + 8 * s, kb * G_BK, n_blk * BN);
            if (p.dense)
              ptx::tma_load_2d(sA + s * G_A_BYTES, &tmap_a, bar_full + 8 * s, kb * G_BK, m_blk * G_BM);
          }
        }
      }
    }
  } else if (!D320 && warp >= G_W_PROD) {
    // ===================== A producers (+ TMA for W) =====================
    // 128 threads; thread = (16-byte chunk c of the 128-byte K-block row, rows rbase + 16 i):
    // 8 consecutive lanes fetch one full 128-byte line -> every cp.async is sector-complete.
    constexpr int RPT = G_BM * 8 / G_PROD_THREADS;       // rows per thread = 8
    constexpr int RSTEP = G_PROD_THREADS / 8;            // 16
    const int pt = (warp - G_W_PROD) * 32 + lane;
    const int c = pt & 7, rbase = pt >> 3;
    uint32_t g = 0;                                      // k-block counter (ring position)
    for (int t = t_begin; t < (p.dense ? t_begin : t_end); t += t_step) {
      const int m_blk = WS ? t : t / p.n_tiles, n_blk = WS ? ws_n_blk : t % p.n_tiles;
      const int m0 = m_blk * G_BM + rbase;
      for (int kb = 0; kb < k_blocks; ++kb, ++g) {
        const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
        const int kglob = kb * G_BK + c * 8;
        const int kk = kglob / p.Cin;
        const int ch = kglob - kk * p.Cin;
        // gather rows for this K-block (issued before the slot wait to overlap its latency)
        int64_t src_row[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int m = m0 + RSTEP * i;
          src_row[i] = m >= p.M ? -1 : (p.idx ? (int64_t)__ldg(p.idx + (size_t)m * p.KD + kk) : (int64_t)m);
        }
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t dst = sA + s * G_A_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = rbase + RSTEP * i;
          const bool ok = src_row[i] >= 0;
          const __nv_bfloat16* src = p.A + (ok ? (size_t)src_row[i] * p.Cin + ch : 0);
          ptx::cp_async16(dst + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        }
        // asynchronous arrive: fires when this thread's copies have landed, so the producer never
        // blocks on data and every free stage of the ring is in flight
        ptx::cp_async_mbar_arrive_noinc(bar_full + 8 * s);
      }
    }
    ptx::cp_async_wait<0>();                             // do not exit with copies in flight
  } else if (warp == G_W_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(G_BM, BN);
      uint32_t g = 0, it = 0;
      if (WS) {
        ptx::mbar_wait(bar_w, 0);
        ptx::tc_fence_after();
      }
      for (int t = t_begin; t < t_end; t += t_step, ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        ptx::mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < k_blocks; ++kb, ++g) {
          const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::fence_proxy_async();                      // cp.async (generic proxy) -> UMMA (async proxy)
          ptx::tc_fence_after();
          const uint64_t ad = ptx::umma_desc_sw128(sA + s * G_A_BYTES);
          const uint64_t bd = ptx::umma_desc_sw128(WS ? sB + kb * (BN * G_BK * 2) : sB + s * G_B_BYTES);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k)
            ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(bar_empty + 8 * s);           // frees the smem stage
        }
        ptx::umma_commit(bar_tfull + 8 * acc);           // accumulator ready
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 0-7) =====================
    // warp % 4 = TMEM lane quadrant (hardware rule), warp / 4 = column half of the tile.
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const int HB = BN >> 1;                              // columns per half (multiple of 32)
    const int cbeg = half * HB;
    const uint32_t stage = sStage + (uint32_t)warp * 2048u;
    float* v_bias = s_vec + half * 384;                  // this half's columns: bias | gamma | beta
    float* v_g = v_bias + 128;
    float* v_b = v_bias + 256;
    const bool do_ln = p.ln_g != nullptr;
    const bool has_res = p.res != nullptr;
    const bool plain_bf16 = WS && !has_res && !do_ln && p.act != 1 && p.out_v_f32 == nullptr && p.out_v_bf16 != nullptr &&
                            p.plain2;
    int loaded_n0 = -1;
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; t += t_step, ++it) {
      const int m_blk = WS ? t : t / p.n_tiles, n_blk = WS ? ws_n_blk : t % p.n_tiles;
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      const int m = m_blk * G_BM + r;
      const int n0 = n_blk * BN;
      if (n0 != loaded_n0) {                             // epilogue vectors -> smem (once per CTA in WS mode)
        asm volatile("bar.sync %0, 128;" ::"r"(5 + half) : "memory");   // previous tile's readers done
        if (r < HB) {
          v_bias[r] = p.bias ? __ldg(p.bias + n0 + cbeg + r) : 0.f;
          if (do_ln) { v_g[r] = __ldg(p.ln_g + cbeg + r); v_b[r] = __ldg(p.ln_b + cbeg + r); }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(5 + half) : "memory");
        loaded_n0 = n0;
      }
      int32_t orow = -1;
      if (m < p.M) orow = p.out_rows ? __ldg(p.out_rows + m) : m;
      const int32_t yrow = p.y_mapped ? orow : (m < p.M ? m : -1);
      uint4 rbuf[8];
      uint4 rbuf2[D320 ? 8 : 1];                         // D320: a second residual chunk in flight
      if (has_res) res_issue(p.res, orow, p.N, n0 + cbeg, lane, rbuf);   // first residual chunk
      if constexpr (D320) {
        if (has_res && HB > 32) res_issue(p.res, orow, p.N, n0 + cbeg + 32, lane, rbuf2);
      }
      ptx::mbar_wait(bar_tfull + 8 * acc, aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + cbeg;
      float shift = 0.f, s1 = 0.f, s2 = 0.f;              // shifted sums for LayerNorm
      uint32_t rawA[32];
      if (plain_bf16) {
        // qkv-type epilogue (+bias, bf16 store): the tensor-memory load of the NEXT chunk is issued as
        // soon as this chunk is packed, so its latency hides behind the transpose / store chain
        auto pack_chunk = [&](const uint32_t (&raw)[32], int col, uint32_t (&w)[16]) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(v_bias + col + 4 * q);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(raw[4 * q]) + bq.x,
                                                      __uint_as_float(raw[4 * q + 1]) + bq.y);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(raw[4 * q + 2]) + bq.z,
                                                      __uint_as_float(raw[4 * q + 3]) + bq.w);
            w[2 * q] = *reinterpret_cast<uint32_t*>(&h0);
            w[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
        };
        char* gb = reinterpret_cast<char*>(p.out_v_bf16);
        ptx::tmem_ld32(taddr, rawA);
        for (int c0 = 0; c0 < HB; c0 += 32) {
          ptx::tmem_ld_wait();
          uint32_t w[16];
          pack_chunk(rawA, c0, w);
          if (c0 + 32 < HB) ptx::tmem_ld32(taddr + c0 + 32, rawA);
          store_unit(stage, lane, w, gb, orow, (size_t)p.N * 2, (size_t)(n0 + cbeg + c0) * 2);
        }
      } else {
      if constexpr (D320) ptx::tmem_ld32(taddr, rawA);
      for (int c0 = 0; c0 < HB; c0 += 32) {
        {
          const int cc = c0;
          uint32_t (&raw32)[32] = rawA;
          if constexpr (!D320) ptx::tmem_ld32(taddr + cc, raw32);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(v_bias + cc + 4 * q);
            v[4 * q] = __uint_as_float(raw32[4 * q]) + b.x;
            v[4 * q + 1] = __uint_as_float(raw32[4 * q + 1]) + b.y;
            v[4 * q + 2] = __uint_as_float(raw32[4 * q + 2]) + b.z;
            v[4 * q + 3] = __uint_as_float(raw32[4 * q + 3]) + b.w;
          }
          if constexpr (D320) {
            // next accumulator chunk under this chunk's work; residual chunks alternate between the two
            // buffers, each refilled two chunks ahead as soon as it has been consumed
            if (cc + 32 < HB) ptx::tmem_ld32(taddr + cc + 32, raw32);
            if (has_res) {
              if ((cc >> 5) & 1) {
                res_add(stage, lane, rbuf2, v);
                if (cc + 64 < HB) res_issue(p.res, orow, p.N, n0 + cbeg + cc + 64, lane, rbuf2);
              } else {
                res_add(stage, lane, rbuf, v);
                if (cc + 64 < HB) res_issue(p.res, orow, p.N, n0 + cbeg + cc + 64, lane, rbuf);
              }
            }
          } else if (has_res) {
            res_add(stage, lane, rbuf, v);
            if (cc + 32 < HB) res_issue(p.res, orow, p.N, n0 + cbeg + cc + 32, lane, rbuf);
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if (p.out_v_f32) store_f32x32(stage, lane, p.out_v_f32, orow, p.N, n0 + cbeg + cc, v);
          if (p.out_v_bf16) store_bf16x32(stage, lane, p.out_v_bf16, orow, p.N, n0 + cbeg + cc, v);
          if (do_ln) {
            if (cc == 0) shift = v[0];
            uint32_t wb[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = v[j] - shift;
              s1 += d; s2 += d * d;
              wb[j] = __float_as_uint(v[j]);
            }
            ptx::tmem_st32(taddr + cc, wb);
          }
        }
      }
      }
      if (do_ln) {
        // per-half (mean, M2) -> exchange with the other column half of this row -> full-row stats
        const float nh = (float)HB;
        const float mean_h = shift + s1 / nh;
        const float m2_h = s2 - s1 * s1 / nh;
        s_ln[half * 128 + r] = make_float2(mean_h, m2_h);
        ptx::tmem_st_wait();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float2 o = s_ln[(half ^ 1) * 128 + r];
        const float delta = o.x - mean_h;
        const float mean = mean_h + 0.5f * delta;
        const float var = (m2_h + o.y + delta * delta * nh * 0.5f) / (2.f * nh);
        const float rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
        for (int c0 = 0; c0 < HB; c0 += 32) {
          {
            const int cc = c0;
            uint32_t (&raw32)[32] = rawA;
            ptx::tmem_ld32(taddr + cc, raw32);
            ptx::tmem_ld_wait();
            float y[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 gm = *reinterpret_cast<const float4*>(v_g + cc + 4 * q);
              const float4 bt = *reinterpret_cast<const float4*>(v_b + cc + 4 * q);
              y[4 * q] = (__uint_as_float(raw32[4 * q]) - mean) * rstd * gm.x + bt.x;
              y[4 * q + 1] = (__uint_as_float(raw32[4 * q + 1]) - mean) * rstd * gm.y + bt.y;
              y[4 * q + 2] = (__uint_as_float(raw32[4 * q + 2]) - mean) * rstd * gm.z + bt.z;
              y[4 * q + 3] = (__uint_as_float(raw32[4 * q + 3]) - mean) * rstd * gm.w + bt.w;
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
            }
            if (p.out_y_f32) store_f32x32(stage, lane, p.out_y_f32, yrow, p.N, cbeg + cc, y);
            if (p.out_y_bf16) store_bf16x32(stage, lane, p.out_y_bf16, yrow, p.N, cbeg + cc, y);
          }
        }
        // the exchange slot is rewritten only after both halves passed the next bar.sync
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == G_W_MMA) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// host side: tensor map for W [N, Ktot] bf16 (K-major), box = 64 x block_n, SW128
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_gather_gemm(const void* A, const int32_t* idx, const void* W, int64_t M, int32_t N,
                    int32_t KD, int32_t Cin, const float* bias, const float* res, int32_t act,
                    float* out_v_f32, void* out_v_bf16, const float* ln_g, const float* ln_b,
                    int32_t relu, int32_t y_mapped, float* out_y_f32, void* out_y_bf16,
                    const int32_t* out_rows, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (M == 0) return HFL_OK;
  HFL_CHECK_ARG(A && W && M > 0 && M < (1ll << 31), "bad A/W/M");
  HFL_CHECK_ARG(N >= 64 && N % 64 == 0 && N <= 4096, "N must be a multiple of 64");
  HFL_CHECK_ARG(KD >= 1 && Cin >= 8 && Cin % 8 == 0, "Cin must be a multiple of 8");
  HFL_CHECK_ARG(((int64_t)KD * Cin) % G_BK == 0, "KD*Cin must be a multiple of 64");
  HFL_CHECK_ARG(Cin % G_BK == 0 || G_BK % Cin == 0, "Cin must divide or be a multiple of 64");
  HFL_CHECK_ARG(idx != nullptr || KD == 1, "KD > 1 needs a gather table");
  const int n_tiles = (N + 255) / 256;
  HFL_CHECK_ARG(N % n_tiles == 0 && (N / n_tiles) % 64 == 0, "N must split into tiles that are multiples of 64");
  const int block_n = N / n_tiles;
  HFL_CHECK_ARG(!ln_g || (n_tiles == 1 && ln_b), "LayerNorm epilogue needs N <= 256");
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled unavailable%s", "");
  CUtensorMap tmap;
  const cuuint64_t Ktot = (cuuint64_t)KD * Cin;
  cuuint64_t dims[2] = {Ktot, (cuuint64_t)N};
  cuuint64_t strides[1] = {Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled failed%s (%lld)", "", (long long)cr);
  // dense Linear (no gather): A tiles are 2-D boxes of the row-major [M, Cin] matrix -> TMA
  // environment toggles (A/B runs, tools/ab_bench.sh) are read once per process
  static const bool env_dense_tma = !(getenv("HFL_GEMM_DENSE_TMA") && getenv("HFL_GEMM_DENSE_TMA")[0] == '0');
  static const bool env_plain2 = !(getenv("HFL_GEMM_PLAIN2") && getenv("HFL_GEMM_PLAIN2")[0] == '0');
  static const bool env_d320 = getenv("HFL_GEMM_DENSE320") && getenv("HFL_GEMM_DENSE320")[0] == '1';
  const int dense = (idx == nullptr && KD == 1 && env_dense_tma);
  CUtensorMap tmap_a = tmap;
  if (dense) {
    cuuint64_t adims[2] = {(cuuint64_t)Cin, (cuuint64_t)M};
    cuuint64_t astr[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t abox[2] = {(cuuint32_t)G_BK, (cuuint32_t)G_BM};
    cr = enc(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), adims, astr, abox, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled (A) failed%s (%lld)", "", (long long)cr);
  }
  GemmParams p;
  p.dense = dense;
  p.plain2 = env_plain2;
  p.A = (const __nv_bfloat16*)A; p.idx = idx; p.M = (int)M; p.N = N; p.KD = KD; p.Cin = Cin;
  p.block_n = block_n; p.n_tiles = n_tiles; p.bias = bias; p.res = res; p.act = act;
  p.out_v_f32 = out_v_f32; p.out_v_bf16 = (__nv_bfloat16*)out_v_bf16; p.ln_g = ln_g; p.ln_b = ln_b;
  p.relu = relu; p.y_mapped = y_mapped; p.out_y_f32 = out_y_f32;
  p.out_y_bf16 = (__nv_bfloat16*)out_y_bf16; p.out_rows = out_rows; p.ln_eps = 1e-5f;
  const int64_t m_tiles = ceil_div(M, G_BM);
  const int sms = sm_count();
  // weight-stationary when the W tile fits next to the A ring and every CTA gets >= 2 M tiles
  p.ws = ((int64_t)Ktot * block_n * 2 <= G_WS_W_BYTES) && (m_tiles * n_tiles >= 2 * sms);
  if (p.ws && dense && env_d320) {          // 320-thread dense instantiation (168 registers; measured: -0.8 ms / step)
    HFL_ENSURE_SMEM(G_SMEM, k_gather_gemm<true, true>);
    const int grid = (sms / n_tiles) * n_tiles;
    HFL_LAUNCH((k_gather_gemm<true, true><<<grid, G_THREADS_D320, G_SMEM, st>>>(tmap, tmap_a, p)));
  } else if (p.ws) {
    HFL_ENSURE_SMEM(G_SMEM, k_gather_gemm<true>);
    const int grid = (sms / n_tiles) * n_tiles;
    HFL_LAUNCH((k_gather_gemm<true><<<grid, G_THREADS, G_SMEM, st>>>(tmap, tmap_a, p)));
  } else {
    HFL_ENSURE_SMEM(G_SMEM, k_gather_gemm<false>);
    const int64_t tiles = m_tiles * n_tiles;
    const int grid = (int)(tiles < sms ? tiles : sms);
    HFL_LAUNCH((k_gather_gemm<false><<<grid, G_THREADS, G_SMEM, st>>>(tmap, tmap_a, p)));
  }
  return HFL_OK;
}

}  // extern "C"
