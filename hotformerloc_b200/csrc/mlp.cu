// Fused transformer MLP on tcgen05:   x += fc2( GELU( fc1(y) + b1 ) ) + b2
// (reference: MLP.forward models/layers/octformer_layers.py:53-59 inside the pre-LN blocks
// octformer_backbone.py:279-281, hotformerloc_backbone.py:215-216, 290-291).
//
// The unfused pair of GEMMs writes and re-reads the (M x 4C) hidden activation through HBM
// (4 KB per token at C = 256 -- a quarter of a block's traffic).  Here one CTA keeps a
// 128-row tile of y resident in smem and walks the hidden dimension in chunks of 128:
//     acc1[b] (TMEM, 128 cols, double buffered) = y_tile . W1[chunk]^T            (UMMA N = 128)
//     epilogue 1: +b1, GELU, bf16 pairs stored back OVER acc1[b]  (H lives in tensor memory)
//     acc2    (TMEM, C cols)                   += H[b] . W2[:, chunk]^T   (UMMA, A operand from TMEM)
// and only the final (128 x C) tile leaves the SM (+b2, +residual, fp32 stream + bf16 shadow).
// Keeping H out of shared memory leaves room for an 8 x 16 KB TMA weight ring (1 MB of weights
// stream from L2 per tile at C = 256).  Warp roles (448 threads, 1 CTA / SM):
//   0-7  epilogue warps (lane quadrant w % 4, column half w / 4): epilogue 1 of every chunk and,
//        between the first two chunks of the next tile, epilogue 2 of the finished tile
//   8    MMA issue + TMEM alloc      9-12  y-tile producers (cp.async)      13  weight TMA producer
// Measured design notes (tools/micro/*.cu): a TMA load costs its issuing thread ~340 ns whatever
// the box size, a lone UMMA N = 256 runs at 100 % / N = 128 at 90 % of the tensor peak, and the
// A-from-TMEM operand layout is lane = row, column c = K elements (2c, 2c + 1).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int ML_BM = 128, ML_CH = 128, ML_STAGE = 16384, ML_SLOT = 32768, ML_ISSUERS = 2;
constexpr int ML_THREADS = 448;

struct MlpParams {
  const __nv_bfloat16* A;     // [M, C] LayerNorm'ed tokens
  int M;
  const float* b1;            // [4C]
  const float* b2;            // [C]
  const float* res;           // fp32 residual stream (row-mapped), also the output
  float* out_f32;
  __nv_bfloat16* out_bf16;    // shadow or NULL
  const int32_t* out_rows;    // [M] or NULL
  int dbg;                    // diagnostics only (HFL_MLP_DBG): 1 no fp32 stores, 2 no bf16 stores, 4 no residual loads
  long long* prof;            // diagnostics only (HFL_MLP_PROF): per-role wait/work cycles of CTA 0
  // PJ variant (attention output projection + residual + LayerNorm in front of the MLP)
  const float* bp;            // [C] proj bias
  const float* ln_g;          // [C] norm2
  const float* ln_b;
  float ln_eps;
  int mc;                     // PJ, C = 256: clusters of two CTAs share every weight load (TMA multicast)
};

// GELU (erf form) of two values at once, given h = x / 2 (the epilogue folds the halving into
// its bias add): GELU = h + t - t * 2^(t P(t)), t = |h| -- the same degree-4 fit as gelu_erf in
// gemm.cu (|err| <= 9e-7), evaluated with packed FFMA2 (8 FP32-pipe ops per element, half the
// issue slots).  Returns the two results rounded and packed as a bf16 pair.
__device__ __forceinline__ uint32_t mlp_gelu2_bf16(ptx::f32x2 h) {
  using namespace ptx;
  const f32x2 t = abs2(h);
  f32x2 q = fma2(bcast2(-1.5639575e-2f), t, bcast2(1.1524727e-1f));
  q = fma2(q, t, bcast2(-4.1725174e-1f));
  q = fma2(q, t, bcast2(-1.8383462f));
  q = fma2(q, t, bcast2(-2.3020070f));
  float a0, a1, e0, e1;
  unpack2(mul2(q, t), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2 y = fma2(t | 0x8000000080000000ull, pack2(e0, e1), add2(h, t));   // (-t) e + (h + t)
  float y0, y1;
  unpack2(y, y0, y1);
  __nv_bfloat162 r = __floats2bfloat162_rn(y0, y1);
  return *reinterpret_cast<uint32_t*>(&r);
}

// ---- coalesced staging helpers (same scheme as gemm.cu) ----
__device__ __forceinline__ uint32_t ml_stage_addr(uint32_t stage, int row, int seg) {
  return stage + (uint32_t)row * 64u + (uint32_t)((seg ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void ml_sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ml_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void ml_store_unit(uint32_t stage, int lane, const uint32_t* w, char* gbase,
                                              int32_t orow, size_t row_bytes, size_t col_byte) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    ml_sts128(ml_stage_addr(stage, lane, q), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = (lane >> 2) + 8 * i, seg = lane & 3;
    const uint4 v = ml_lds128(ml_stage_addr(stage, row, seg));
    const int32_t orr = __shfl_sync(0xffffffffu, orow, row);
    if (orr >= 0) *reinterpret_cast<uint4*>(gbase + (size_t)orr * row_bytes + col_byte + seg * 16) = v;
  }
}
__device__ __forceinline__ void ml_res_issue(const float* base, int32_t orow, int ld, int col, int lane,
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
__device__ __forceinline__ void ml_res_add(uint32_t stage, int lane, const uint4 (&buf)[8], float (&v)[32]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (lane >> 2) + 8 * i, seg = lane & 3;
      const uint4 b = buf[u * 4 + i];
      ml_sts128(ml_stage_addr(stage, row, seg), b.x, b.y, b.z, b.w);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 b = ml_lds128(ml_stage_addr(stage, lane, q));
      v[u * 16 + 4 * q] += __uint_as_float(b.x);
      v[u * 16 + 4 * q + 1] += __uint_as_float(b.y);
      v[u * 16 + 4 * q + 2] += __uint_as_float(b.z);
      v[u * 16 + 4 * q + 3] += __uint_as_float(b.w);
    }
  }
}

// Fast path of epilogue 2 for the block kernel (PJ, contiguous output rows): acc2 already holds
// s + b2 + fc2(...), so there is no bias and no residual to add and the output row of a tile row is
// arithmetic (no shuffles).  One call = the 32 accumulator columns from col0 of the warp's lane quadrant.
// Per 16-column unit the rows go through the warp's transposing stage (lane l then holds 4 consecutive
// columns of rows (l >> 2) + 8 i: 64 B contiguous per 4 lanes); the four transposed loads of a unit are
// issued together and its global stores run behind the stage traffic of the next unit.  The stores are the
// slow part (the SM's write path, ~30 B / clock): the block kernel gives them to warps of their own.
template <int C>
__device__ __forceinline__ void ml_epi2_slice(const MlpParams& p, int tile, int quad, int lane, uint32_t lane_base,
                                              uint32_t stage, int col0, uint32_t arrive_bar) {
  const int row0 = tile * ML_BM + quad * 32 + (lane >> 2), seg = lane & 3;
  uint32_t raw[32];
  ptx::tmem_ld32(lane_base + 256 + col0, raw);
  ptx::tmem_ld_wait();
  if (arrive_bar) {                                          // acc2 fully read by this warp
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(arrive_bar);
  }
  uint4 a[4];
  auto flush = [&](int c0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int orr = row0 + 8 * i;
      if (orr < p.M) {
        const size_t col = (size_t)(c0 + seg * 4);
        if (!(p.dbg & 1)) *reinterpret_cast<uint4*>(p.out_f32 + (size_t)orr * C + col) = a[i];
        if (p.out_bf16 && !(p.dbg & 2)) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(a[i].x), __uint_as_float(a[i].y));
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(a[i].z), __uint_as_float(a[i].w));
          *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)orr * C + col) =
              make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
      }
    }
  };
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    __syncwarp();                                            // the previous unit's transposed loads have completed
#pragma unroll
    for (int qd = 0; qd < 4; ++qd)
      ml_sts128(ml_stage_addr(stage, lane, qd), raw[u * 16 + 4 * qd], raw[u * 16 + 4 * qd + 1],
                raw[u * 16 + 4 * qd + 2], raw[u * 16 + 4 * qd + 3]);
    if (u == 1) flush(col0);                                 // global stores of unit 0
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = ml_lds128(ml_stage_addr(stage, (lane >> 2) + 8 * i, seg));
  }
  flush(col0 + 16);
  __syncwarp();                                              // the stage is free again
}

// Same for the store warps of the block kernel, in 32-column units: the 2 KB stage takes 16 rows x 128 B at a time
// (16-byte chunk c of row r at c ^ (r & 7)); after the transposition 8 lanes hold one row's 128 contiguous bytes,
// so every fp32 store instruction writes FOUR FULL LINES (the 16-column form writes eight half lines: twice the
// wavefronts on the SM's store path, which is what bounds this epilogue) and every bf16 store four 64-byte pieces.
template <int C>
__device__ __forceinline__ void ml_epi2_slice32(const MlpParams& p, int tile, int quad, int lane, uint32_t stage,
                                                int col0, const uint32_t (&raw)[32]) {
  const int ch = lane & 7;
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {                     // rows 0-15, then 16-31 of the quadrant
    __syncwarp();                                            // the previous pass's transposed loads have completed
    if ((lane >> 4) == hrow) {
      const int lr = lane & 15;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        ml_sts128(stage + (uint32_t)lr * 128u + (uint32_t)((c ^ (lr & 7)) << 4), raw[4 * c], raw[4 * c + 1],
                  raw[4 * c + 2], raw[4 * c + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = (lane >> 3) + 4 * i;                    // row of this pass, 8 lanes per row
      const uint4 a = ml_lds128(stage + (uint32_t)rr * 128u + (uint32_t)((ch ^ (rr & 7)) << 4));
      const int orr = tile * ML_BM + quad * 32 + hrow * 16 + rr;
      if (orr < p.M) {
        const size_t col = (size_t)(col0 + ch * 4);
        if (!(p.dbg & 1)) *reinterpret_cast<uint4*>(p.out_f32 + (size_t)orr * C + col) = a;
        if (p.out_bf16 && !(p.dbg & 2)) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(a.x), __uint_as_float(a.y));
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(a.z), __uint_as_float(a.w));
          *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)orr * C + col) =
              make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
      }
    }
  }
  __syncwarp();
}

template <int C, bool PJ = false>
struct MlpSmem {
  static constexpr int KB1 = C / 64;                 // K blocks of GEMM 1
  static constexpr int NCH = 4 * C / ML_CH;          // hidden chunks
  static constexpr int W2S = C / 128;                // 16 KB stages per W2 K block (N2 = C rows)
  static constexpr int RING = 4;                     // 32 KB weight slots (one hidden chunk of W1 + W2 at C = 256)
  static constexpr int G1S = KB1 / 2;                // slots GEMM 1 consumes per chunk (two K blocks per slot)
  static constexpr int G2S = W2S;                    // slots GEMM 2 consumes per chunk
  static constexpr int A1_BYTES = KB1 * ML_STAGE;
  static constexpr int W_BYTES = RING * ML_SLOT;
  static constexpr int ST_BYTES = 8 * 2048;          // one transposing stage per epilogue warp
  static constexpr int OFF_W = A1_BYTES;
  static constexpr int OFF_ST = OFF_W + W_BYTES;
  static constexpr int OFF_B1 = OFF_ST + ST_BYTES;   // b1 [4C] fp32
  static constexpr int OFF_B2 = OFF_B1 + 4 * C * 4;  // b2 [C] fp32
  static constexpr int OFF_BAR = OFF_B2 + (PJ ? 0 : C * 4);   // PJ: b2 is folded into epilogue 0 (s_pj)
  static constexpr int OFF_PJ = OFF_BAR + 256;       // PJ: proj bias | norm2 gamma | norm2 beta | b2   [4][C] fp32
  static constexpr int OFF_LNX = OFF_PJ + 4 * C * 4; // PJ: LayerNorm statistics exchange [2 halves][128 rows] float2
  static constexpr int OFF_ST2 = OFF_LNX + 2048;     // PJ: transposing stages of the four store warps
  static constexpr int TOTAL = PJ ? OFF_ST2 + 4 * 2048 : OFF_BAR + 256;
};

// cycle accounting for CTA 0 (diagnostics; prof == nullptr in production)
#define ML_T0() const long long t0_ = PROF ? clock64() : 0
#define ML_ACC(slot) do { if (PROF) lacc[slot] += clock64() - t0_; } while (0)

// PJ = true: the attention output projection, its residual add and the block's norm2 run in front of
// the MLP inside the same CTA (reference: x = x + proj(attn) ; x = x + mlp(norm2(x)),
// octformer_backbone.py:276-281, hotformerloc_backbone.py:212-216):
//     acc1 region (C cols)  = o_tile . Wp^T                         GEMM 0, o tile by TMA into the A buffer
//     epilogue 0: s = acc + bp + residual (+ b2) -> acc2 (fp32, TENSOR MEMORY); LayerNorm(s - b2) -> bf16
//                 straight into the swizzled A buffer (the MLP operand: no HBM round trip of `y`)
//     GEMM 1 / epilogue 1 / GEMM 2 as before, GEMM 2 accumulating ON TOP of s
//     epilogue 2: acc2 -> x (fp32) + bf16 shadow: the residual stream is read once per block here
template <int C, bool PROF, bool PJ>
__global__ void __launch_bounds__(ML_THREADS, 1)
k_mlp_fused(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
            const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_wp,
            const MlpParams p) {
  using S = MlpSmem<C, PJ>;
  constexpr int KB1 = S::KB1, NCH = S::NCH, RING = S::RING, G1S = S::G1S, G2S = S::G2S;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ptx::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sA1 = base, sW = base + S::OFF_W, sSt = base + S::OFF_ST;
  float* s_b1 = reinterpret_cast<float*>(smem + S::OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + S::OFF_B2);
  const uint32_t bar = base + S::OFF_BAR;
  const uint32_t w_full = bar, w_empty = bar + 64;            // RING + RING
  const uint32_t a1_full = bar + 128, a1_empty = bar + 160;   // KB1 (<= 4), 1
  const uint32_t c1_full = bar + 168;                         // acc1[2]: GEMM 1 done
  const uint32_t h_full = bar + 184;                          // H[2] (aliases acc1): epilogue 1 done
  const uint32_t c2_full = bar + 200, c2_empty = bar + 208;   // acc2
  const uint32_t s_tmem = bar + 216;
  const uint32_t o_full = bar + 224, c0_full = bar + 232;     // PJ: o tile landed, GEMM 0 done
  const uint32_t seed_done = bar + 240;                       // PJ split mode: s + b2 moved into acc2
  // PJ with contiguous output rows (the block kernel): the four otherwise idle warps 9-12 write the finished tile
  // out (epilogue 2) while warps 0-7 already run epilogue 0 of the next tile
  const bool split = PJ && !p.out_rows && !(p.dbg & 8);
  float* s_pj = reinterpret_cast<float*>(smem + S::OFF_PJ);   // PJ: bp | gamma | beta | b2
  float2* s_lnx = reinterpret_cast<float2*>(smem + S::OFF_LNX);
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + S::OFF_BAR + 216);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + ML_BM - 1) / ML_BM;
  // Weight multicast (mc): the two CTAs of a cluster stream the SAME weights in the same order -- every slot is
  // filled by two half-size loads, one issued by each CTA and delivered to both (half the L2 reads and half the
  // TMA issue work per SM; the chunk phase demands ~60 B/clock of weights per SM, about what L2 gives 148
  // independent streams).  Both run the same number of tiles (a dummy tile beyond M stores nothing).
  const bool mc = PJ && p.mc != 0;
  const uint32_t crank = mc ? ptx::cluster_ctarank() : 0u;
  int n_my = (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  if (mc) {
    const int n_peer = (m_tiles - (int)(blockIdx.x ^ 1u) + (int)gridDim.x - 1) / (int)gridDim.x;
    n_my = n_my > n_peer ? n_my : n_peer;
  }
  // every CTA (pair) walks the hidden chunks in its own rotation (spreads the L2 reads)
  const int rot = (mc ? (blockIdx.x >> 1) : blockIdx.x) % NCH;
  long long lacc[23];                 // per-thread cycle counters (registers; dead code unless profiling)
#pragma unroll
  for (int i = 0; i < 23; ++i) lacc[i] = 0;
  const long long t_role = PROF ? clock64() : 0;

  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) s_b1[i] = 0.5f * p.b1[i];   // epilogue 1 works on (acc + b1) / 2
  if constexpr (!PJ) for (int i = threadIdx.x; i < C; i += blockDim.x) s_b2[i] = p.b2[i];
  if constexpr (PJ) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      s_pj[i] = p.bp[i]; s_pj[C + i] = p.ln_g[i]; s_pj[2 * C + i] = p.ln_b[i]; s_pj[3 * C + i] = p.b2[i];
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) { ptx::mbar_init(w_full + 8 * s, 1); ptx::mbar_init(w_empty + 8 * s, mc ? 2 : 1); }
    for (int k = 0; k < 4; ++k) ptx::mbar_init(a1_full + 8 * k, PJ ? 4 : 128);   // PJ: the 4 epilogue warps of the K block's column half
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(c0_full, 1);
    ptx::mbar_init(a1_empty, 1);
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(c1_full + 8 * b, 1); ptx::mbar_init(h_full + 8 * b, 8); }
    ptx::mbar_init(c2_full, 1);
    ptx::mbar_init(c2_empty, split ? 4 : 8);
    ptx::mbar_init(seed_done, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 8) { ptx::tmem_alloc(s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (mc) ptx::cluster_sync();                               // the peer's barriers exist before anything is sent to them
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t t_acc2 = tmem_base + 256;

  if (warp == 13) {
    // ===================== weight TMA producer =====================
    // A TMA load costs its issuing thread ~340 ns whatever the box size
    // (tools/micro/tma_issue.cu), so the stages are dealt round-robin to ML_ISSUERS lanes.
    if (lane < ML_ISSUERS) {
      ptx::prefetch_tmap(&tm_w1);
      ptx::prefetch_tmap(&tm_w2);
      uint32_t g = 0;
      // every load is one 32 KB box (64 K-columns, rows, K blocks): {64, 128, 2} or {64, 256, 1}
      // mc: the tensor maps carry HALF boxes ({64, 128, 1}); by_rows = the slot's halves are row halves of one K
      // block (W2, Wp), else the two K blocks of the slot (W1)
      auto load = [&](const CUtensorMap* tm, int row0, int kblk, bool by_rows) {
        if ((int)(g % (uint32_t)ML_ISSUERS) == lane) {
          const uint32_t s = g % RING, ph = (g / RING) & 1;
          // mc: w_empty collects the MMA completion of BOTH CTAs (multicast commit): the slot is free in the pair
          { ML_T0(); ptx::mbar_wait_sleep(w_empty + 8 * s, ph ^ 1, 64); ML_ACC(13); }
          if (mc) {
            ptx::mbar_arrive_expect_tx(w_full + 8 * s, ML_SLOT);
            ptx::tma_load_3d_mc(sW + s * ML_SLOT + crank * (ML_SLOT / 2), tm, w_full + 8 * s, 0,
                                by_rows ? row0 + (int)crank * 128 : row0, by_rows ? kblk : kblk + (int)crank, 0x3);
          } else {
            ptx::mbar_arrive_expect_tx(w_full + 8 * s, ML_SLOT);
            ptx::tma_load_3d(sW + s * ML_SLOT, tm, w_full + 8 * s, 0, row0, kblk);
          }
        }
        ++g;
      };
      auto load_w1 = [&](int j) {
        const int jr = (j + rot) % NCH;
        for (int h = 0; h < G1S; ++h) load(&tm_w1, jr * ML_CH, 2 * h, false);
      };
      auto load_w2 = [&](int j) {
        const int jr = (j + rot) % NCH;
        if (C == 256) { load(&tm_w2, 0, 2 * jr, true); load(&tm_w2, 0, 2 * jr + 1, true); }
        else load(&tm_w2, 0, 2 * jr, true);
      };
      // same order as the MMA issuer: G1(0) G1(1) | G2(j) G1(j+2) ... (G1 runs into the next tile)
      for (int it = 0; it < n_my; ++it) {
        if constexpr (PJ) {
          // Wp (K-major [C, C]): C = 256 -> four {64, 256, 1} boxes, C = 128 -> one {64, 128, 2} box
          if (C == 256) { for (int kb = 0; kb < 4; ++kb) load(&tm_wp, 0, kb, true); }
          else load(&tm_wp, 0, 0, true);
          load_w1(0); load_w1(1);
          for (int j = 0; j < NCH; ++j) {
            load_w2(j);
            if (j + 2 < NCH) load_w1(j + 2);
          }
        } else {
          if (it == 0) { load_w1(0); load_w1(1); }
          for (int j = 0; j < NCH; ++j) {
            load_w2(j);
            if (j + 2 < NCH) load_w1(j + 2);
            else if (it + 1 < n_my) load_w1(j + 2 - NCH);
          }
        }
      }
      if (PROF) lacc[14] = clock64() - t_role;
    } else if (PJ && lane == ML_ISSUERS) {
      // o tiles (attention output, the A operand of GEMM 0): four / two 128 x 64 boxes into the A buffer,
      // as soon as GEMM 1 of the previous tile's last chunk has released it.  Own lane: the weight
      // issuers never wait on the A buffer.
      ptx::prefetch_tmap(&tm_o);
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        ptx::mbar_wait_sleep(a1_empty, (it & 1) ^ 1, 128);
        ptx::mbar_arrive_expect_tx(o_full, KB1 * ML_STAGE);
        for (int kb = 0; kb < KB1; ++kb)
          ptx::tma_load_2d(sA1 + kb * ML_STAGE, &tm_o, o_full, kb * 64, tile * ML_BM);
      }
    }
  } else if (warp >= 9) {
    // ===================== y-tile producers =====================
    if (split) {
      // ===================== store warps (block kernel): epilogue 2 of every tile =====================
      const int q2 = warp & 3;                               // TMEM lane quadrant of this warp (hardware: warp % 4)
      const uint32_t lb2 = tmem_base + ((uint32_t)(q2 * 32) << 16);
      const uint32_t st2 = base + S::OFF_ST2 + (uint32_t)(warp - 9) * 2048u;
      for (int it = 0; it < n_my; ++it) {
        ptx::mbar_wait_sleep(c2_full, it & 1, 64);
        ptx::tc_fence_after();
        // the next 32-column slice is in flight from tensor memory while the current one is transposed and stored
        uint32_t raw[2][32];
        ptx::tmem_ld32(lb2 + 256, raw[0]);
#pragma unroll
        for (int sl = 0; sl < C / 32; ++sl) {
          ptx::tmem_ld_wait();
          if (sl + 1 < C / 32) ptx::tmem_ld32(lb2 + 256 + (sl + 1) * 32, raw[(sl + 1) & 1]);
          else {                                             // acc2 fully read by this warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(c2_empty);
          }
          ml_epi2_slice32<C>(p, blockIdx.x + it * gridDim.x, q2, lane, st2, sl * 32, raw[sl & 1]);
        }
      }
    }
    const int pt = (warp - 9) * 32 + lane;
    const int c = pt & 7, rbase = pt >> 3;
    for (int it = 0; it < (PJ ? 0 : n_my); ++it) {       // PJ: the A buffer is filled by TMA (o) and by epilogue 0 (y)
      const int tile = blockIdx.x + it * gridDim.x;
      { ML_T0(); ptx::mbar_wait_sleep(a1_empty, (it & 1) ^ 1, 256); ML_ACC(15); }
      const int m0 = tile * ML_BM + rbase;
#pragma unroll
      for (int kb = 0; kb < KB1; ++kb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rbase + 16 * i, m = m0 + 16 * i;
          const bool ok = m < p.M;
          const __nv_bfloat16* src = p.A + (ok ? (size_t)m * C + kb * 64 + c * 8 : 0);
          ptx::cp_async16(sA1 + kb * ML_STAGE + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4), src,
                          ok ? 16u : 0u);
        }
        ptx::cp_async_mbar_arrive_noinc(a1_full + 8 * kb);
      }
    }
    ptx::cp_async_wait<0>();
    if (PROF) lacc[16] = clock64() - t_role;
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16(ML_BM, ML_CH);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(ML_BM, C);
      uint32_t g = 0;
      // GEMM 1 of chunk j of this CTA's tile number t:  acc1[q & 1] = y_tile . W1[chunk]^T
      auto gemm1 = [&](int t, int j) {
        const uint32_t q = (uint32_t)t * NCH + j, b = q & 1;
        for (int h = 0; h < G1S; ++h, ++g) {
          if (j == 0) { ML_T0(); ptx::mbar_wait(a1_full + 8 * (2 * h), t & 1); ptx::mbar_wait(a1_full + 8 * (2 * h + 1), t & 1);
                        if (split && h == 0) ptx::mbar_wait(seed_done, t & 1); ML_ACC(0); }
          const uint32_t s = g % RING, ph = (g / RING) & 1;
          { ML_T0(); ptx::mbar_wait(w_full + 8 * s, ph); ML_ACC(1); }
          ML_T0();
          ptx::fence_proxy_async();                          // the y tile was written with cp.async
          ptx::tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t ad = ptx::umma_desc_sw128(sA1 + (2 * h + kk) * ML_STAGE);
            const uint64_t bd = ptx::umma_desc_sw128(sW + s * ML_SLOT + kk * ML_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(tmem_base + b * ML_CH, ad + 2 * k, bd + 2 * k, idesc1, (h | kk | k) != 0);
          }
          if (mc) ptx::umma_commit_mc(w_empty + 8 * s, 0x3); else ptx::umma_commit(w_empty + 8 * s);
          ML_ACC(5);
        }
        ptx::umma_commit(c1_full + 8 * b);
        if (j == NCH - 1) ptx::umma_commit(a1_empty);
      };
      // GEMM 2:  acc2 += H[q & 1] . W2[:, chunk]^T with H read from TENSOR MEMORY (bf16 pairs that
      // epilogue 1 packed over the fp32 accumulator it had just read; K 0-63 at columns 0-31,
      // K 64-127 at columns 64-95 of acc1[b])
      auto gemm2 = [&](int t, int j) {
        const uint32_t q = (uint32_t)t * NCH + j, b = q & 1;
        { ML_T0(); ptx::mbar_wait(h_full + 8 * b, (q >> 1) & 1); ML_ACC(2); }
        if (j == 0) { ML_T0(); ptx::mbar_wait(c2_empty, (t & 1) ^ 1); ML_ACC(3); }
        ptx::tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          // C = 256: one slot per K block (256 rows x 64);  C = 128: both K blocks in one slot
          const bool fresh = C == 256 || kb == 0;
          const uint32_t s = g % RING, ph = (g / RING) & 1;
          if (fresh) { ML_T0(); ptx::mbar_wait(w_full + 8 * s, ph); ML_ACC(4); }
          ML_T0();
          ptx::tc_fence_after();
          const uint32_t ta = tmem_base + b * ML_CH + kb * 64;
          const uint64_t bd = ptx::umma_desc_sw128(sW + s * ML_SLOT + (C == 256 ? 0 : kb * ML_STAGE));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ts(t_acc2, ta + 8 * k, bd + 2 * k, idesc2, PJ || (j | kb | k) != 0);   // PJ: acc2 was seeded by epilogue 0
          if (C == 256 || kb == 1) { if (mc) ptx::umma_commit_mc(w_empty + 8 * s, 0x3); else ptx::umma_commit(w_empty + 8 * s); ++g; }
          ML_ACC(6);
        }
        if (j == NCH - 1) ptx::umma_commit(c2_full);
      };
      // The tensor pipe executes one thread's MMAs in issue order, so GEMM 1 of chunk q + 2 (which
      // overwrites acc1[q & 1]) is simply issued after GEMM 2 of chunk q (which reads H from it).
      // GEMM 0 (PJ): acc1 region = o_tile . Wp^T; it may start while epilogue 2 of the previous tile still
      // reads acc2 (different columns); the H operands it overwrites were consumed by GEMM 2 in issue order
      auto gemm0 = [&](int t) {
        const uint32_t idesc0 = ptx::umma_idesc_bf16(ML_BM, C);
        ptx::mbar_wait(o_full, t & 1);
        for (int h = 0; h < (C == 256 ? 4 : 1); ++h, ++g) {
          const uint32_t s = g % RING, ph = (g / RING) & 1;
          ptx::mbar_wait(w_full + 8 * s, ph);
          ptx::tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < (C == 256 ? 1 : 2); ++kk) {
            const int kb = C == 256 ? h : kk;
            const uint64_t ad = ptx::umma_desc_sw128(sA1 + kb * ML_STAGE);
            const uint64_t bd = ptx::umma_desc_sw128(sW + s * ML_SLOT + (C == 256 ? 0 : kk * ML_STAGE));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc0, (kb | k) != 0);
          }
          if (mc) ptx::umma_commit_mc(w_empty + 8 * s, 0x3); else ptx::umma_commit(w_empty + 8 * s);
        }
        ptx::umma_commit(c0_full);
      };
      for (int it = 0; it < n_my; ++it) {
        if constexpr (PJ) {
          gemm0(it);
          gemm1(it, 0); gemm1(it, 1);
          for (int j = 0; j < NCH; ++j) {
            gemm2(it, j);
            if (j + 2 < NCH) gemm1(it, j + 2);
          }
        } else {
          if (it == 0) { gemm1(0, 0); gemm1(0, 1); }
          for (int j = 0; j < NCH; ++j) {
            gemm2(it, j);
            if (j + 2 < NCH) gemm1(it, j + 2);
            else if (it + 1 < n_my) gemm1(it + 1, j + 2 - NCH);
          }
        }
      }
      if (PROF) lacc[7] = clock64() - t_role;
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps 0-7 =====================
    // warp w owns TMEM lane quadrant w % 4 (rows) and column half w / 4 of whatever it reads.
    const int quad = warp & 3, half = warp >> 2, r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t stage = sSt + (uint32_t)warp * 2048u;

    // epilogue 1: acc1 -> +b1, GELU -> bf16 pairs written back over the accumulator (H in TMEM)
    auto epi1 = [&](int t, int j) {
      const uint32_t q = (uint32_t)t * NCH + j, b = q & 1;
      const int jr = (j + rot) % NCH;
      { ML_T0(); ptx::mbar_wait(c1_full + 8 * b, (q >> 1) & 1); ML_ACC(8); }
      ptx::tc_fence_after();
      ML_T0();
      const uint32_t tacc = lane_base + b * ML_CH + half * 64;
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        uint32_t raw[32];
        long long tq0 = PROF ? clock64() : 0;
        ptx::tmem_ld32(tacc + sl * 32, raw);
        ptx::tmem_ld_wait();
        if (PROF) { const long long tq = clock64(); lacc[20] += tq - tq0; tq0 = tq; }
        uint32_t w[16];
        const float* bias = s_b1 + jr * ML_CH + half * 64 + sl * 32;
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * qd);
          const ptx::f32x2 hf = ptx::bcast2(0.5f);
          w[2 * qd] = mlp_gelu2_bf16(ptx::fma2(ptx::pack2u(raw[4 * qd], raw[4 * qd + 1]), hf, ptx::pack2(bb.x, bb.y)));
          w[2 * qd + 1] = mlp_gelu2_bf16(ptx::fma2(ptx::pack2u(raw[4 * qd + 2], raw[4 * qd + 3]), hf, ptx::pack2(bb.z, bb.w)));
        }
        if (PROF) { const long long tq = clock64(); lacc[21] += tq - tq0; tq0 = tq; }
        ptx::tmem_st16(tacc + sl * 16, w);                   // always behind this warp's own reads
        if (PROF) { lacc[22] += clock64() - tq0; }
      }
      { const long long tq0 = PROF ? clock64() : 0;
      ptx::tmem_st_wait();
      if (PROF) lacc[22] += clock64() - tq0; }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(h_full + 8 * b);
      ML_ACC(9);
    };

    // epilogue 2: acc2 + b2 + residual -> x (fp32) [+ bf16 shadow], coalesced through the stage
    // one 32-column slice of this warp's column half of tile t (see ml_epi2_slice)
    auto epi2_slice = [&](int t, int sl, bool last) {
      ml_epi2_slice<C>(p, blockIdx.x + t * gridDim.x, quad, lane, lane_base, stage, half * (C / 2) + sl * 32,
                       last ? c2_empty : 0u);
    };
    auto epi2_fast = [&](int t) {
      { ML_T0(); ptx::mbar_wait(c2_full, t & 1); ML_ACC(10); }
      ptx::tc_fence_after();
      ML_T0();
#pragma unroll 1
      for (int sl = 0; sl < C / 64; ++sl) epi2_slice(t, sl, sl + 1 == C / 64);
      ML_ACC(11);
    };
    auto epi2 = [&](int t) {
      if (PJ && !p.out_rows && !(p.dbg & 8)) { epi2_fast(t); return; }
      const int tile = blockIdx.x + t * gridDim.x;
      const int m = tile * ML_BM + r;
      int32_t orow = -1;
      if (m < p.M) orow = p.out_rows ? __ldg(p.out_rows + m) : m;
      constexpr int CW = C / 2;                              // columns of this warp
      const int cbase = half * CW;
      uint4 rbuf[8];
      ml_res_issue(p.res, (PJ || (p.dbg & 4)) ? -1 : orow, C, cbase, lane, rbuf);   // PJ: residual already inside acc2
      { ML_T0(); ptx::mbar_wait(c2_full, t & 1); ML_ACC(10); }
      ptx::tc_fence_after();
      ML_T0();
      const uint32_t taddr = lane_base + 256 + cbase;
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        uint32_t raw[32];
        long long tq0 = PROF ? clock64() : 0;
        ptx::tmem_ld32(taddr + c0, raw);
        ptx::tmem_ld_wait();
        if (PROF) { const long long tq = clock64(); lacc[17] += tq - tq0; tq0 = tq; }
        if (c0 + 32 >= CW) {                                 // acc2 fully read: GEMM 2 of the next tile may start
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(c2_empty);
        }
        // Per 16-column unit: transpose acc2 + b2 through the warp's stage so that lane l holds 4
        // consecutive columns (seg = l & 3) of rows (l >> 2) + 8 i -- the layout the residual was
        // loaded in -- then add and store from there: fp32 as 16 B, the bf16 shadow as 8 B per lane.
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          __syncwarp();
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const float4 bb = PJ ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(s_b2 + cbase + c0 + u * 16 + 4 * qd);
            ml_sts128(ml_stage_addr(stage, lane, qd), __float_as_uint(__uint_as_float(raw[u * 16 + 4 * qd]) + bb.x),
                      __float_as_uint(__uint_as_float(raw[u * 16 + 4 * qd + 1]) + bb.y),
                      __float_as_uint(__uint_as_float(raw[u * 16 + 4 * qd + 2]) + bb.z),
                      __float_as_uint(__uint_as_float(raw[u * 16 + 4 * qd + 3]) + bb.w));
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = (lane >> 2) + 8 * i, seg = lane & 3;
            const uint4 a = ml_lds128(ml_stage_addr(stage, row, seg));
            const uint4 rr = rbuf[u * 4 + i];
            const float o0 = __uint_as_float(a.x) + __uint_as_float(rr.x), o1 = __uint_as_float(a.y) + __uint_as_float(rr.y);
            const float o2 = __uint_as_float(a.z) + __uint_as_float(rr.z), o3 = __uint_as_float(a.w) + __uint_as_float(rr.w);
            const int32_t orr = __shfl_sync(0xffffffffu, orow, row);
            if (orr >= 0) {
              const size_t col = (size_t)(cbase + c0 + u * 16 + seg * 4);
              if (!(p.dbg & 1))
                *reinterpret_cast<float4*>(p.out_f32 + (size_t)orr * C + col) = make_float4(o0, o1, o2, o3);
              if (p.out_bf16 && !(p.dbg & 2)) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
                *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)orr * C + col) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
              }
            }
            // this register is free again: fetch the same piece of the next 32-column slice
            if (c0 + 32 < CW)
              rbuf[u * 4 + i] = (!PJ && orr >= 0 && !(p.dbg & 4))
                                    ? *reinterpret_cast<const uint4*>(p.res + (size_t)orr * C + cbase + c0 + 32 + u * 16 + seg * 4)
                                    : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        if (PROF) lacc[19] += clock64() - tq0;
      }
      ML_ACC(11);
    };

    // epilogue 0 (PJ): s = o.Wp^T + bp + residual; acc2 <- s + b2 (the seed of GEMM 2); A buffer <- LayerNorm(s)
    auto epi0 = [&](int t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int m = tile * ML_BM + r;
      int32_t orow = -1;
      if (m < p.M) orow = p.out_rows ? __ldg(p.out_rows + m) : m;
      constexpr int CW = C / 2;
      const int cbase = half * CW;
      const float* v_bp = s_pj + cbase;
      const float* v_g = s_pj + C + cbase;
      const float* v_be = s_pj + 2 * C + cbase;
      const float* v_b2 = s_pj + 3 * C + cbase;
      uint4 rbuf[8];
      ml_res_issue(p.res, orow, C, cbase, lane, rbuf);
      { ML_T0(); ptx::mbar_wait(c0_full, t & 1); ML_ACC(15); }
      ptx::tc_fence_after();
      const long long t_e0 = PROF ? clock64() : 0;
      const uint32_t t_src = lane_base + cbase, t_dst = lane_base + 256 + cbase;
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld32(t_src + c0, raw);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(v_bp + c0 + 4 * q);
          v[4 * q] = __uint_as_float(raw[4 * q]) + b.x;
          v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b.y;
          v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b.z;
          v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b.w;
        }
        ml_res_add(stage, lane, rbuf, v);
        if (c0 + 32 < CW) ml_res_issue(p.res, orow, C, cbase + c0 + 32, lane, rbuf);
        if (c0 == 0) shift = v[0];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(v_b2 + c0 + 4 * q);
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = v[4 * q + e] - shift;
            s1 += d; s2 = fmaf(d, d, s2);
            raw[4 * q + e] = __float_as_uint(v[4 * q + e] + (split ? 0.f : bb[e]));
          }
        }
        ptx::tmem_st32((split ? t_src : t_dst) + c0, raw);
      }
      // per-half (mean, M2) -> exchange with the warp that owns the other column half of this row
      const float nh = (float)CW;
      const float mean_h = shift + s1 / nh;
      const float m2_h = s2 - s1 * s1 / nh;
      s_lnx[half * 128 + r] = make_float2(mean_h, m2_h);
      ptx::tmem_st_wait();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      const float2 o = s_lnx[(half ^ 1) * 128 + r];
      const float delta = o.x - mean_h;
      const float mean = mean_h + 0.5f * delta;
      const float var = (m2_h + o.y + delta * delta * nh * 0.5f) / (2.f * nh);
      const float rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld32((split ? t_src : t_dst) + c0, raw);
        ptx::tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 gm = *reinterpret_cast<const float4*>(v_g + c0 + 4 * q);
          const float4 bt = *reinterpret_cast<const float4*>(v_be + c0 + 4 * q);
          float4 b2 = *reinterpret_cast<const float4*>(v_b2 + c0 + 4 * q);
          if (split) b2 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float y0 = (__uint_as_float(raw[4 * q]) - b2.x - mean) * rstd * gm.x + bt.x;
          const float y1 = (__uint_as_float(raw[4 * q + 1]) - b2.y - mean) * rstd * gm.y + bt.y;
          const float y2 = (__uint_as_float(raw[4 * q + 2]) - b2.z - mean) * rstd * gm.z + bt.z;
          const float y3 = (__uint_as_float(raw[4 * q + 3]) - b2.w - mean) * rstd * gm.w + bt.w;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(y2, y3);
          w[2 * q] = *reinterpret_cast<uint32_t*>(&h0);
          w[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        // row r of the K-major 128B-swizzled operand: K block = 64 columns, 16-byte chunk c at (c ^ (r & 7))
        const int col = cbase + c0, kb = col >> 6, ch0 = (col & 63) >> 3;
        const uint32_t rowa = sA1 + kb * ML_STAGE + (uint32_t)r * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          ml_sts128(rowa + (uint32_t)(((ch0 + q) ^ (r & 7)) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        if (((col + 32) & 63) == 0) {                        // this warp's share of K block kb is complete
          ptx::fence_proxy_async();                          // generic-proxy stores -> UMMA (async proxy)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(a1_full + 8 * kb);
        }
      }
      // the exchange slot is rewritten only after both halves passed this barrier
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      if (split) {
        // seed of GEMM 2: s + b2 from the GEMM-0 columns into acc2, as soon as the store warps have read the
        // previous tile out of it; GEMM 1 of chunk 0 (which overwrites these columns) waits for seed_done
        if (t > 0) { ML_T0(); ptx::mbar_wait(c2_empty, (t & 1) ^ 1); ML_ACC(3); }
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < CW; c0 += 32) {
          uint32_t raw[32];
          ptx::tmem_ld32(t_src + c0, raw);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(v_b2 + c0 + 4 * q);
            raw[4 * q] = __float_as_uint(__uint_as_float(raw[4 * q]) + b.x);
            raw[4 * q + 1] = __float_as_uint(__uint_as_float(raw[4 * q + 1]) + b.y);
            raw[4 * q + 2] = __float_as_uint(__uint_as_float(raw[4 * q + 2]) + b.z);
            raw[4 * q + 3] = __float_as_uint(__uint_as_float(raw[4 * q + 3]) + b.w);
          }
          ptx::tmem_st32(t_dst + c0, raw);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(seed_done);
      }
      if (PROF) lacc[18] += clock64() - t_e0;
    };

    // static schedule: the final tile of iteration t - 1 is written out between the first two
    // chunks of iteration t, when its GEMM 2 has long finished and acc2 is about to be reused
    // (PJ: before epilogue 0 of iteration t, which re-seeds acc2)
    for (int it = 0; it < n_my; ++it) {
      {                                                      // pull this tile's residual rows (of this warp's column half) into L2; they are read one tile later
        const int m = (blockIdx.x + it * gridDim.x) * ML_BM + r;
        if (m < p.M) {
          const int32_t orow = p.out_rows ? __ldg(p.out_rows + m) : m;
          const char* rp = reinterpret_cast<const char*>(p.res + (size_t)orow * C + half * (C / 2));
#pragma unroll
          for (int k = 0; k < C * 2 / 128; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + k * 128));
        }
      }
      if constexpr (PJ) {
        if (it > 0 && !split) epi2(it - 1);                  // split: the store warps write the tiles out
        epi0(it);
        for (int j = 0; j < NCH; ++j) epi1(it, j);
      } else {
        epi1(it, 0);
        if (it > 0) epi2(it - 1);
        for (int j = 1; j < NCH; ++j) epi1(it, j);
      }
    }
    if (n_my > 0 && !split) epi2(n_my - 1);
    if (PROF) lacc[12] = clock64() - t_role;
  }
  if (PROF && blockIdx.x == 0) {
    auto dump = [&](int a, int b) { for (int i = a; i < b; ++i) p.prof[i] = lacc[i]; };
    if (threadIdx.x == 8 * 32) dump(0, 8);
    if (threadIdx.x == 0) { dump(8, 13); dump(17, 23); if (PJ) dump(15, 16); }
    if (threadIdx.x == 13 * 32) dump(13, 15);
    if (threadIdx.x == 9 * 32 && !PJ) dump(15, 17);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (mc) ptx::cluster_sync();                               // nothing of the peer is in flight towards this CTA any more
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled2 mlp_get_encode() {
  static PFN_encodeTiled2 fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled2)q;
  }
  return fn;
}

template <int C, bool PROF, bool PJ>
static int launch_mlp(const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& to, const CUtensorMap& tp,
                      const MlpParams& p, cudaStream_t st) {
  // the attribute is per device: set it on every launch (cheap) instead of caching a process-wide flag
  HFL_CUDA(cudaFuncSetAttribute(k_mlp_fused<C, PROF, PJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MlpSmem<C, PJ>::TOTAL));
  const int m_tiles = (p.M + ML_BM - 1) / ML_BM;
  const int sms = sm_count();
  int grid = m_tiles < sms ? m_tiles : sms;
  if (p.mc) {
    // clusters of two CTAs (weight multicast): an even grid, launched with the cluster attribute
    grid = (grid + 1) & ~1;
    if (grid > (sms & ~1)) grid = sms & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ML_THREADS); cfg.dynamicSmemBytes = MlpSmem<C, PJ>::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    HFL_CUDA(cudaLaunchKernelEx(&cfg, k_mlp_fused<C, PROF, PJ>, t1, t2, to, tp, p));
    ::hfl::g_launches.fetch_add(1, std::memory_order_relaxed);
    return HFL_OK;
  }
  HFL_LAUNCH((k_mlp_fused<C, PROF, PJ><<<grid, ML_THREADS, MlpSmem<C, PJ>::TOTAL, st>>>(t1, t2, to, tp, p)));
  return HFL_OK;
}

static int mlp_dispatch(const void* A, const void* Wp, const float* bp, const float* ln_g, const float* ln_b,
                        const void* W1, const float* b1, const void* W2, const float* b2, int64_t M, int32_t C,
                        const float* res, float* out_f32, void* out_bf16, const int32_t* out_rows,
                        cudaStream_t st) {
  const bool pj = Wp != nullptr;
  PFN_encodeTiled2 enc = mlp_get_encode();
  if (!enc) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled unavailable%s", "");
  CUtensorMap t1, t2, to, tp;
  // 3-D views {64 K-columns, rows, K blocks of 64}: one box = consecutive 16 KB K-major UMMA tiles
  auto make_map = [&](CUtensorMap* tm, const void* W, int rows, int cols, int box_rows, int box_kb) -> CUresult {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_kb}, es[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(W), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  // weight multicast between the two CTAs of a cluster (block kernel, C = 256): half boxes, see the kernel
  static const bool env_mc = !(getenv("HFL_MLP_MULTICAST") && getenv("HFL_MLP_MULTICAST")[0] == '0');
  const bool mc = pj && C == 256 && env_mc && M >= 2 * ML_BM;
  CUresult cr = mc ? make_map(&t1, W1, 4 * C, C, 128, 1) : make_map(&t1, W1, 4 * C, C, 128, 2);
  if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (W1) failed%s (%lld)", "", (long long)cr);
  cr = mc ? make_map(&t2, W2, C, 4 * C, 128, 1) : C == 256 ? make_map(&t2, W2, C, 4 * C, 256, 1) : make_map(&t2, W2, C, 4 * C, 128, 2);
  if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (W2) failed%s (%lld)", "", (long long)cr);
  to = t1; tp = t1;
  if (pj) {
    cr = mc ? make_map(&tp, Wp, C, C, 128, 1) : C == 256 ? make_map(&tp, Wp, C, C, 256, 1) : make_map(&tp, Wp, C, C, 128, 2);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (Wp) failed%s (%lld)", "", (long long)cr);
    // o tile: 128 x 64 boxes of the row-major [M, C] matrix, rows >= M zero-filled
    cuuint64_t adims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    cuuint64_t astr[1] = {(cuuint64_t)C * 2};
    cuuint32_t abox[2] = {64, (cuuint32_t)ML_BM}, es[2] = {1, 1};
    cr = enc(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), adims, astr, abox, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (o) failed%s (%lld)", "", (long long)cr);
  }
  static long long* prof = nullptr;
  static const bool want_prof = getenv("HFL_MLP_PROF") != nullptr;
  static const int dbg = getenv("HFL_MLP_DBG") ? atoi(getenv("HFL_MLP_DBG")) : 0;
  if (want_prof && !prof) cudaMalloc(&prof, 32 * sizeof(long long));
  if (want_prof) cudaMemsetAsync(prof, 0, 32 * sizeof(long long), st);
  MlpParams p{(const __nv_bfloat16*)A, (int)M, b1, b2, res, out_f32, (__nv_bfloat16*)out_bf16, out_rows,
              dbg, want_prof ? prof : nullptr, bp, ln_g, ln_b, 1e-5f, mc ? 1 : 0};
  int rc;
  if (pj && want_prof && C == 256) rc = launch_mlp<256, true, true>(t1, t2, to, tp, p, st);
  else if (pj) rc = C == 128 ? launch_mlp<128, false, true>(t1, t2, to, tp, p, st) : launch_mlp<256, false, true>(t1, t2, to, tp, p, st);
  else if (want_prof) rc = C == 128 ? launch_mlp<128, true, false>(t1, t2, to, tp, p, st) : launch_mlp<256, true, false>(t1, t2, to, tp, p, st);
  else rc = C == 128 ? launch_mlp<128, false, false>(t1, t2, to, tp, p, st) : launch_mlp<256, false, false>(t1, t2, to, tp, p, st);
  if (want_prof && (!pj || C == 256) && rc == HFL_OK) {
    long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* nm[] = {"mma:a1_full", "mma:w_full(g1)", "mma:h_full", "mma:c2_empty", "mma:w_full(g2)",
                               "mma:issue(g1)", "mma:issue(g2)", "mma:total", "epi:c1_full", "epi:e1_work",
                               "epi:c2_full", "epi:e2_work", "epi:total", "tma:w_empty", "tma:total",
                               "y:a1_empty|e0:c0_full", "y:total", "e2:tmem_ld", "e0:work", "e2:transpose+stores", "e1:tmem_ld", "e1:gelu", "e1:tmem_st"};
    fprintf(stderr, "[hfl_%smlp_fused prof C=%d M=%lld tiles/CTA=%.1f]", pj ? "proj_" : "", C, (long long)M, (double)((M + 127) / 128) / sm_count());
    for (int i = 0; i < 23; ++i) fprintf(stderr, " %s=%.1fk", nm[i], h[i] / 1e3);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_mlp_fused(const void* A, const void* W1, const float* b1, const void* W2, const float* b2,
                  int64_t M, int32_t C, const float* res, float* out_f32, void* out_bf16,
                  const int32_t* out_rows, void* stream_) {
  if (M == 0) return HFL_OK;
  HFL_CHECK_ARG(A && W1 && b1 && W2 && b2 && res && out_f32, "null argument");
  HFL_CHECK_ARG(C == 128 || C == 256, "C must be 128 or 256");
  HFL_CHECK_ARG(M > 0 && M < (1ll << 31), "bad M");
  return mlp_dispatch(A, nullptr, nullptr, nullptr, nullptr, W1, b1, W2, b2, M, C, res, out_f32, out_bf16,
                      out_rows, (cudaStream_t)stream_);
}

int hfl_proj_mlp_fused(const void* O, const void* Wp, const float* bp, const float* ln_g, const float* ln_b,
                       const void* W1, const float* b1, const void* W2, const float* b2, int64_t M, int32_t C,
                       const float* res, float* out_f32, void* out_bf16, const int32_t* out_rows,
                       void* stream_) {
  if (M == 0) return HFL_OK;
  HFL_CHECK_ARG(O && Wp && bp && ln_g && ln_b && W1 && b1 && W2 && b2 && res && out_f32, "null argument");
  HFL_CHECK_ARG(C == 128 || C == 256, "C must be 128 or 256");
  HFL_CHECK_ARG(M > 0 && M < (1ll << 31), "bad M");
  return mlp_dispatch(O, Wp, bp, ln_g, ln_b, W1, b1, W2, b2, M, C, res, out_f32, out_bf16, out_rows,
                      (cudaStream_t)stream_);
}

}  // extern "C"
