// Fused transformer MLP on tcgen05:   x += fc2( GELU( fc1(y) + b1 ) ) + b2
// (reference: MLP.forward models/layers/octformer_layers.py:53-59 inside the pre-LN blocks
// octformer_backbone.py:279-281, hotformerloc_backbone.py:215-216, 290-291).
//
// The unfused pair of GEMMs writes and re-reads the (M x 4C) hidden activation through HBM
// (4 KB per token at C = 256 -- a quarter of a block's traffic).  Here one CTA keeps a
// 128-row tile of y resident in smem and walks the hidden dimension in chunks of 128:
//     acc1[b] (TMEM, 128 cols, double buffered) = y_tile . W1[chunk]^T            (UMMA N = 128)
//     epilogue-1 warps: +b1, GELU, bf16 -> smem H[b] in the UMMA K-major swizzled layout
//     acc2    (TMEM, C cols)                   += H[b] . W2[:, chunk]^T            (UMMA N = C)
// and only the final (128 x C) tile leaves the SM (+b2, +residual, fp32 stream + bf16 shadow).
// Weights stream from L2 through a 6 x 16 KB TMA ring (1 MB per tile at C = 256: the kernel is
// bound by that L2 stream, not by HBM).  Warp roles (448 threads, 1 CTA / SM):
//   0-3  epilogue 1 (one TMEM lane quadrant each)      8      MMA issue + TMEM alloc
//   4-7  epilogue 2 (final tile, coalesced I/O)        9-12   y-tile producers (cp.async)
//                                                      13     weight TMA producer
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int ML_BM = 128, ML_CH = 128, ML_RING = 6, ML_STAGE = 16384;
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
  int dbg;                    // diagnostics only (HFL_MLP_DBG): 1 no GELU, 2 no H stores, 4 no proxy fence in MMA
};

__device__ __forceinline__ float mlp_gelu(float x) {
  // same folded-constant erf form as the GEMM epilogue (|err| <= 2.6e-6)
  const float t = fabsf(x), s = x * x;
  float r = fmaf(-4.40836608e-6f, t, 1.38209148e-4f);
  const float u = fmaf(-9.90546318e-4f, t, 8.74800568e-3f);
  r = fmaf(r, s * 0.5f, u);
  r = fmaf(r, t, -5.44641622e-2f);
  r = fmaf(r, t, -4.57945084e-1f);
  r = fmaf(r, t, -1.15144926f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(r * t));
  const float ht = 0.5f * t;
  return fmaf(-ht, e, fmaf(0.5f, x, ht));
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

template <int C>
struct MlpSmem {
  static constexpr int KB1 = C / 64;                 // K blocks of GEMM 1
  static constexpr int NCH = 4 * C / ML_CH;          // hidden chunks
  static constexpr int W2S = C / 128;                // 16 KB stages per W2 K block (N2 = C rows)
  static constexpr int A1_BYTES = KB1 * ML_STAGE;
  static constexpr int H_BYTES = 32768;              // single buffer: the weight ring needs the smem
  static constexpr int W_BYTES = ML_RING * ML_STAGE;
  static constexpr int ST_BYTES = 4 * 2048;
  static constexpr int OFF_H = A1_BYTES;
  static constexpr int OFF_W = OFF_H + H_BYTES;
  static constexpr int OFF_ST = OFF_W + W_BYTES;
  static constexpr int OFF_B1 = OFF_ST + ST_BYTES;   // b1 [4C] fp32
  static constexpr int OFF_B2 = OFF_B1 + 4 * C * 4;  // b2 [C] fp32
  static constexpr int OFF_BAR = OFF_B2 + C * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
};

template <int C>
__global__ void __launch_bounds__(ML_THREADS, 1)
k_mlp_fused(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
            const MlpParams p) {
  using S = MlpSmem<C>;
  constexpr int KB1 = S::KB1, NCH = S::NCH, W2S = S::W2S;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ptx::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sA1 = base, sH = base + S::OFF_H, sW = base + S::OFF_W, sSt = base + S::OFF_ST;
  float* s_b1 = reinterpret_cast<float*>(smem + S::OFF_B1);
  float* s_b2 = reinterpret_cast<float*>(smem + S::OFF_B2);
  const uint32_t bar = base + S::OFF_BAR;
  const uint32_t w_full = bar, w_empty = bar + 64;            // 8 + 8 slots (ML_RING used)
  const uint32_t a1_full = bar + 128, a1_empty = bar + 160;   // 4, 1
  const uint32_t c1_full = bar + 168, c1_empty = bar + 184;   // acc1: 2 + 2
  const uint32_t h_full = bar + 200, h_empty = bar + 208;     // 1 + 1 (single H buffer)
  const uint32_t c2_full = bar + 216, c2_empty = bar + 224;   // acc2: 1 + 1
  const uint32_t s_tmem = bar + 232;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + S::OFF_BAR + 232);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + ML_BM - 1) / ML_BM;

  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) s_b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_b2[i] = p.b2[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < ML_RING; ++s) { ptx::mbar_init(w_full + 8 * s, 1); ptx::mbar_init(w_empty + 8 * s, 1); }
    for (int k = 0; k < 4; ++k) ptx::mbar_init(a1_full + 8 * k, 128);
    ptx::mbar_init(a1_empty, 1);
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(c1_full + 8 * b, 1); ptx::mbar_init(c1_empty + 8 * b, 4); }
    ptx::mbar_init(h_full, 4);
    ptx::mbar_init(h_empty, 1);
    ptx::mbar_init(c2_full, 1);
    ptx::mbar_init(c2_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 8) { ptx::tmem_alloc(s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t t_acc2 = tmem_base + 256;

  if (warp == 13) {
    // ===================== weight TMA producer =====================
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_w1);
      ptx::prefetch_tmap(&tm_w2);
      uint32_t g = 0;
      auto load = [&](const CUtensorMap* tm, int c0, int c1) {
        const uint32_t s = g % ML_RING, ph = (g / ML_RING) & 1;
        ptx::mbar_wait(w_empty + 8 * s, ph ^ 1);
        ptx::mbar_arrive_expect_tx(w_full + 8 * s, ML_STAGE);
        ptx::tma_load_2d(sW + s * ML_STAGE, tm, w_full + 8 * s, c0, c1);
        ++g;
      };
      auto load_w1 = [&](int j) { for (int kb = 0; kb < KB1; ++kb) load(&tm_w1, kb * 64, j * ML_CH); };
      auto load_w2 = [&](int j) {
        for (int kb = 0; kb < 2; ++kb)
          for (int hf = 0; hf < W2S; ++hf) load(&tm_w2, j * ML_CH + kb * 64, hf * 128);
      };
      // every CTA walks the hidden chunks in a different rotation so that the 148 CTAs do not
      // hammer the same L2 lines in lockstep
      const int rot = blockIdx.x % NCH;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        load_w1(rot);
        for (int j = 0; j < NCH; ++j) {
          if (j + 1 < NCH) load_w1((j + 1 + rot) % NCH);
          load_w2((j + rot) % NCH);
        }
      }
    }
  } else if (warp >= 9) {
    // ===================== y-tile producers =====================
    const int pt = (warp - 9) * 32 + lane;
    const int c = pt & 7, rbase = pt >> 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      ptx::mbar_wait(a1_empty, (it & 1) ^ 1);
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
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16(ML_BM, ML_CH);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(ML_BM, C);
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        auto gemm1 = [&](int j) {
          const int b = j & 1;
          const uint32_t use = it * (NCH / 2) + (j >> 1);
          ptx::mbar_wait(c1_empty + 8 * b, (use & 1) ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < KB1; ++kb, ++g) {
            if (j == 0) ptx::mbar_wait(a1_full + 8 * kb, it & 1);
            const uint32_t s = g % ML_RING, ph = (g / ML_RING) & 1;
            ptx::mbar_wait(w_full + 8 * s, ph);
            if (!(p.dbg & 4)) ptx::fence_proxy_async();
            ptx::tc_fence_after();
            const uint64_t ad = ptx::umma_desc_sw128(sA1 + kb * ML_STAGE);
            const uint64_t bd = ptx::umma_desc_sw128(sW + s * ML_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(tmem_base + b * ML_CH, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0);
            ptx::umma_commit(w_empty + 8 * s);
          }
          ptx::umma_commit(c1_full + 8 * b);
          if (j == NCH - 1) ptx::umma_commit(a1_empty);
        };
        auto gemm2 = [&](int j) {
          const uint32_t huse = it * NCH + j;
          ptx::mbar_wait(h_full, huse & 1);
          if (j == 0) ptx::mbar_wait(c2_empty, (it & 1) ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t s0 = g % ML_RING;
            for (int hf = 0; hf < W2S; ++hf) {
              const uint32_t s = (g + hf) % ML_RING, ph = ((g + hf) / ML_RING) & 1;
              ptx::mbar_wait(w_full + 8 * s, ph);
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_after();
            const uint64_t ad = ptx::umma_desc_sw128(sH + kb * ML_STAGE);
            const uint64_t bd = ptx::umma_desc_sw128(sW + s0 * ML_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(t_acc2, ad + 2 * k, bd + 2 * k, idesc2, (j | kb | k) != 0);
            for (int hf = 0; hf < W2S; ++hf) ptx::umma_commit(w_empty + 8 * ((g + hf) % ML_RING));
            g += W2S;
          }
          ptx::umma_commit(h_empty);
          if (j == NCH - 1) ptx::umma_commit(c2_full);
        };
        gemm1(0);
        for (int j = 0; j < NCH; ++j) {
          if (j + 1 < NCH) gemm1(j + 1);
          gemm2(j);
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ===================== epilogue 1: acc1 -> +b1, GELU -> H (smem, UMMA layout) =====================
    const int quad = warp, r = quad * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      for (int j = 0; j < NCH; ++j) {
        const int b = j & 1;
        const uint32_t use = it * (NCH / 2) + (j >> 1);
        const uint32_t huse = it * NCH + j;
        const int jr = (j + (int)(blockIdx.x % NCH)) % NCH;    // rotated chunk (see the TMA producer)
        ptx::mbar_wait(c1_full + 8 * b, use & 1);
        ptx::tc_fence_after();
        ptx::mbar_wait(h_empty, (huse & 1) ^ 1);             // GEMM 2 of the previous chunk has read H
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + b * ML_CH;
        const uint32_t hrow = sH + (uint32_t)r * 128u;
#pragma unroll 1
        for (int c0 = 0; c0 < ML_CH; c0 += 32) {
          uint32_t raw[32];
          ptx::tmem_ld32(taddr + c0, raw);
          ptx::tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = *reinterpret_cast<const float4*>(s_b1 + jr * ML_CH + c0 + 4 * q);
            float v0 = __uint_as_float(raw[4 * q]) + bb.x, v1 = __uint_as_float(raw[4 * q + 1]) + bb.y;
            float v2 = __uint_as_float(raw[4 * q + 2]) + bb.z, v3 = __uint_as_float(raw[4 * q + 3]) + bb.w;
            if (!(p.dbg & 1)) { v0 = mlp_gelu(v0); v1 = mlp_gelu(v1); v2 = mlp_gelu(v2); v3 = mlp_gelu(v3); }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            w[2 * q] = *reinterpret_cast<uint32_t*>(&h0);
            w[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
          const uint32_t kbase = hrow + (uint32_t)(c0 >> 6) * ML_STAGE;
          const int ch0 = (c0 & 63) >> 3;                      // first 16-byte chunk of these 32 columns
          if (!(p.dbg & 2))
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            ml_sts128(kbase + (uint32_t)(((ch0 + q4) ^ (r & 7)) << 4), w[4 * q4], w[4 * q4 + 1],
                      w[4 * q4 + 2], w[4 * q4 + 3]);
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async();                            // H written by the generic proxy, read by UMMA
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(h_full); ptx::mbar_arrive(c1_empty + 8 * b); }
      }
    }
  } else {
    // ===================== epilogue 2: acc2 + b2 + residual -> x (fp32) [+ bf16 shadow] =====================
    const int quad = warp - 4, r = quad * 32 + lane;
    const uint32_t stage = sSt + (uint32_t)quad * 2048u;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const int m = tile * ML_BM + r;
      int32_t orow = -1;
      if (m < p.M) orow = p.out_rows ? __ldg(p.out_rows + m) : m;
      uint4 rbuf[8];
      ml_res_issue(p.res, orow, C, 0, lane, rbuf);
      ptx::mbar_wait(c2_full, it & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = t_acc2 + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld32(taddr + c0, raw);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = *reinterpret_cast<const float4*>(s_b2 + c0 + 4 * q);
          v[4 * q] = __uint_as_float(raw[4 * q]) + bb.x;
          v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bb.y;
          v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bb.z;
          v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bb.w;
        }
        ml_res_add(stage, lane, rbuf, v);
        if (c0 + 32 < C) ml_res_issue(p.res, orow, C, c0 + 32, lane, rbuf);
        uint32_t w[16];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int k = 0; k < 16; ++k) w[k] = __float_as_uint(v[u * 16 + k]);
          ml_store_unit(stage, lane, w, reinterpret_cast<char*>(p.out_f32), orow, (size_t)C * 4,
                        (size_t)(c0 + u * 16) * 4);
        }
        if (p.out_bf16) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
          }
          ml_store_unit(stage, lane, w, reinterpret_cast<char*>(p.out_bf16), orow, (size_t)C * 2, (size_t)c0 * 2);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(c2_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
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

template <int C>
static int launch_mlp(const CUtensorMap& t1, const CUtensorMap& t2, const MlpParams& p, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    HFL_CUDA(cudaFuncSetAttribute(k_mlp_fused<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpSmem<C>::TOTAL));
    attr = true;
  }
  const int m_tiles = (p.M + ML_BM - 1) / ML_BM;
  const int grid = m_tiles < kSMs ? m_tiles : kSMs;
  HFL_LAUNCH((k_mlp_fused<C><<<grid, ML_THREADS, MlpSmem<C>::TOTAL, st>>>(t1, t2, p)));
  return HFL_OK;
}

}  // namespace hfl

using namespace hfl;

extern "C" {

int hfl_mlp_fused(const void* A, const void* W1, const float* b1, const void* W2, const float* b2,
                  int64_t M, int32_t C, const float* res, float* out_f32, void* out_bf16,
                  const int32_t* out_rows, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (M == 0) return HFL_OK;
  HFL_CHECK_ARG(A && W1 && b1 && W2 && b2 && res && out_f32, "null argument");
  HFL_CHECK_ARG(C == 128 || C == 256, "C must be 128 or 256");
  HFL_CHECK_ARG(M > 0 && M < (1ll << 31), "bad M");
  PFN_encodeTiled2 enc = mlp_get_encode();
  if (!enc) return fail(HFL_ERR_CUDA, "cuTensorMapEncodeTiled unavailable%s", "");
  CUtensorMap t1, t2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)4 * C};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult cr = enc(&t1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W1), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (W1) failed%s (%lld)", "", (long long)cr);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)4 * C, (cuuint64_t)C};
    cuuint64_t strides[1] = {(cuuint64_t)4 * C * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult cr = enc(&t2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W2), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(HFL_ERR_CUDA, "tensor map (W2) failed%s (%lld)", "", (long long)cr);
  }
  const char* dbg_env = getenv("HFL_MLP_DBG");
  MlpParams p{(const __nv_bfloat16*)A, (int)M, b1, b2, res, out_f32, (__nv_bfloat16*)out_bf16, out_rows,
              dbg_env ? atoi(dbg_env) : 0};
  return C == 128 ? launch_mlp<128>(t1, t2, p, st) : launch_mlp<256>(t1, t2, p, st);
}

}  // extern "C"
