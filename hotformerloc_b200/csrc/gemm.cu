// Gather-GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), the workhorse of the
// dense part of the hot path:
//
//   out[m, :] = epilogue( sum_{kk < KD} A[idx[m, kk], :] . W[:, kk*Cin : (kk+1)*Cin]^T )
//
// * KD == 1, idx == NULL  : plain token-major Linear (qkv / proj / fc1 / fc2 / mixer)
// * KD == 27 or 8         : ocnn OctreeConv as an implicit GEMM -- the im2col buffer of
//                           the reference (rows x kdim x Cin) is never materialised;
//                           rows are gathered straight into the swizzled smem operand.
// Persistent, warp-specialised CTA (288 threads, 1 CTA / SM):
//   warps 0-3  epilogue   (TMEM lane quadrant = warp id; thread == output row)
//   warp  4    TMEM alloc + single-thread tcgen05.mma issue (M=128, N=block_n, K=16)
//   warps 5-8  A producers: thread == tile row, 8 x 16 B cp.async (zero-fill for
//              neigh < 0 / tail rows) into the 128B-swizzled K-major layout;
//              first producer thread also issues the TMA load of the weight tile.
// 4-stage smem ring (A 16 KB + B <= 32 KB per stage), double-buffered TMEM accumulators
// (2 x block_n columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue fusions: +bias, +residual (fp32 stream), GELU, LayerNorm over the full row
// (thread-local: one thread owns one row of TMEM), ReLU, fp32 and/or bf16 stores with an
// optional output row map (relay-token rows / hierarchical window layout).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_STAGES = 4;
constexpr int G_LAG = 2;
constexpr int G_A_BYTES = G_BM * G_BK * 2;    // 16 KB
constexpr int G_B_BYTES = 256 * G_BK * 2;     // 32 KB (block_n <= 256)
constexpr int G_THREADS = 288;
constexpr int G_SMEM = G_STAGES * (G_A_BYTES + G_B_BYTES) + 256 + 1024;

struct GemmParams {
  const __nv_bfloat16* A;   // [rows_A, Cin]
  const int32_t* idx;       // [M, KD] row gather table or NULL (identity, KD == 1)
  int M, N, KD, Cin;        // Ktot = KD * Cin
  int block_n, n_tiles;
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

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 p = __floats2bfloat162_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&p);
    }
    d4[q] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

__global__ void __launch_bounds__(G_THREADS, 1)
k_gather_gemm(const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = base + G_STAGES * G_A_BYTES;
  const uint32_t sBar = sB + G_STAGES * G_B_BYTES;
  const uint32_t bar_full = sBar;                       // G_STAGES x 8 B
  const uint32_t bar_empty = sBar + 8 * G_STAGES;       // G_STAGES x 8 B
  const uint32_t bar_tfull = sBar + 16 * G_STAGES;      // 2 x 8 B
  const uint32_t bar_tempty = bar_tfull + 16;           // 2 x 8 B
  const uint32_t s_tmem = bar_tempty + 16;              // 4 B
  volatile uint32_t* tmem_ptr_s =
      reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + G_BM - 1) / G_BM;
  const int total_tiles = m_tiles * p.n_tiles;
  const int k_blocks = (p.KD * p.Cin) / G_BK;
  const int BN = p.block_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 128 + 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(s_tmem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp >= 5) {
    // ===================== A producers (+ TMA for W) =====================
    const int r = (warp - 5) * 32 + lane;                // tile row
    const bool tma_thread = (warp == 5 && lane == 0);
    if (tma_thread) ptx::prefetch_tmap(&tmap_w);
    const uint32_t b_bytes = (uint32_t)BN * G_BK * 2;
    uint32_t g = 0;                                      // k-block counter (ring position)
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m_blk = t / p.n_tiles, n_blk = t % p.n_tiles;
      const int m = m_blk * G_BM + r;
      const bool row_ok = m < p.M;
      const int32_t* idx_row = (p.idx && row_ok) ? p.idx + (size_t)m * p.KD : nullptr;
      int cached_kk = -1;
      int64_t cached_src = -1;
      for (int kb = 0; kb < k_blocks; ++kb, ++g) {
        const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
        ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (tma_thread) {
          ptx::mbar_arrive_expect_tx(bar_full + 8 * s, b_bytes);
          ptx::tma_load_2d(sB + s * G_B_BYTES, &tmap_w, bar_full + 8 * s, kb * G_BK, n_blk * BN);
        }
        const uint32_t dst_row = sA + s * G_A_BYTES + (uint32_t)r * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int kglob = kb * G_BK + c * 8;
          const int kk = kglob / p.Cin;
          const int ch = kglob - kk * p.Cin;
          if (kk != cached_kk) {
            cached_kk = kk;
            cached_src = !row_ok ? -1 : (idx_row ? (int64_t)__ldg(idx_row + kk) : (int64_t)m);
          }
          const bool ok = cached_src >= 0;
          const __nv_bfloat16* src = p.A + (ok ? (size_t)cached_src * p.Cin + ch : 0);
          ptx::cp_async16(dst_row + (uint32_t)((c ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        }
        ptx::cp_async_commit();
        if (g >= G_LAG) {
          ptx::cp_async_wait<G_LAG>();
          ptx::fence_proxy_async();
          ptx::mbar_arrive(bar_full + 8 * ((g - G_LAG) % G_STAGES));
        }
      }
    }
    ptx::cp_async_wait<0>();
    ptx::fence_proxy_async();
    const uint32_t first = g >= G_LAG ? g - G_LAG : 0;
    for (uint32_t q = first; q < g; ++q) ptx::mbar_arrive(bar_full + 8 * (q % G_STAGES));
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(G_BM, BN);
      uint32_t g = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        ptx::mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < k_blocks; ++kb, ++g) {
          const uint32_t s = g % G_STAGES, ph = (g / G_STAGES) & 1;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          ptx::tc_fence_after();
          const uint64_t ad = ptx::umma_desc_sw128(sA + s * G_A_BYTES);
          const uint64_t bd = ptx::umma_desc_sw128(sB + s * G_B_BYTES);
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
    // ===================== epilogue (warps 0-3) =====================
    const int r = warp * 32 + lane;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int m_blk = t / p.n_tiles, n_blk = t % p.n_tiles;
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      const int m = m_blk * G_BM + r;
      const int n0 = n_blk * BN;
      ptx::mbar_wait(bar_tfull + 8 * acc, aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 256;
      int64_t orow = -1;
      if (m < p.M) orow = p.out_rows ? (int64_t)__ldg(p.out_rows + m) : (int64_t)m;
      const bool live = orow >= 0;
      const int64_t yrow = p.y_mapped ? orow : (int64_t)m;
      const bool do_ln = p.ln_g != nullptr;
      float sum = 0.f;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t raw32[32];
        ptx::tmem_ld32(taddr + c0, raw32);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw32[j]);
        if (p.bias) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0) + q);
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
          }
        }
        if (p.res && live) {
          const float4* rp = reinterpret_cast<const float4*>(p.res + orow * p.N + n0 + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 b = rp[q];
            v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (live) {
          if (p.out_v_f32) store_f32x32(p.out_v_f32 + orow * p.N + n0 + c0, v);
          if (p.out_v_bf16) store_bf16x32(p.out_v_bf16 + orow * p.N + n0 + c0, v);
        }
        if (do_ln) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { sum += v[j]; raw32[j] = __float_as_uint(v[j]); }
          ptx::tmem_st32(taddr + c0, raw32);
        }
      }
      if (do_ln) {
        ptx::tmem_st_wait();
        const float mean = sum / (float)BN;
        float ss = 0.f;
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw32[32];
          ptx::tmem_ld32(taddr + c0, raw32);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { float d = __uint_as_float(raw32[j]) - mean; ss += d * d; }
        }
        const float rstd = rsqrtf(ss / (float)BN + p.ln_eps);
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw32[32];
          ptx::tmem_ld32(taddr + c0, raw32);
          ptx::tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 gm = __ldg(reinterpret_cast<const float4*>(p.ln_g + c0) + q);
            float4 bt = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0) + q);
            y[4 * q] = (__uint_as_float(raw32[4 * q]) - mean) * rstd * gm.x + bt.x;
            y[4 * q + 1] = (__uint_as_float(raw32[4 * q + 1]) - mean) * rstd * gm.y + bt.y;
            y[4 * q + 2] = (__uint_as_float(raw32[4 * q + 2]) - mean) * rstd * gm.z + bt.z;
            y[4 * q + 3] = (__uint_as_float(raw32[4 * q + 3]) - mean) * rstd * gm.w + bt.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
          }
          if (yrow >= 0 && m < p.M) {
            if (p.out_y_f32) store_f32x32(p.out_y_f32 + yrow * p.N + c0, y);
            if (p.out_y_bf16) store_bf16x32(p.out_y_bf16 + yrow * p.N + c0, y);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
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
  HFL_CHECK_ARG(N >= 32 && N % 32 == 0 && N <= 4096, "N must be a multiple of 32");
  HFL_CHECK_ARG(KD >= 1 && Cin >= 8 && Cin % 8 == 0, "Cin must be a multiple of 8");
  HFL_CHECK_ARG(((int64_t)KD * Cin) % G_BK == 0, "KD*Cin must be a multiple of 64");
  HFL_CHECK_ARG(Cin % G_BK == 0 || G_BK % Cin == 0, "Cin must divide or be a multiple of 64");
  HFL_CHECK_ARG(idx != nullptr || KD == 1, "KD > 1 needs a gather table");
  const int n_tiles = (N + 255) / 256;
  HFL_CHECK_ARG(N % n_tiles == 0 && (N / n_tiles) % 32 == 0, "N not tileable");
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
  GemmParams p;
  p.A = (const __nv_bfloat16*)A; p.idx = idx; p.M = (int)M; p.N = N; p.KD = KD; p.Cin = Cin;
  p.block_n = block_n; p.n_tiles = n_tiles; p.bias = bias; p.res = res; p.act = act;
  p.out_v_f32 = out_v_f32; p.out_v_bf16 = (__nv_bfloat16*)out_v_bf16; p.ln_g = ln_g; p.ln_b = ln_b;
  p.relu = relu; p.y_mapped = y_mapped; p.out_y_f32 = out_y_f32;
  p.out_y_bf16 = (__nv_bfloat16*)out_y_bf16; p.out_rows = out_rows; p.ln_eps = 1e-5f;
  static bool attr_set = false;
  if (!attr_set) {
    HFL_CUDA(cudaFuncSetAttribute(k_gather_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    attr_set = true;
  }
  const int64_t tiles = ceil_div(M, G_BM) * n_tiles;
  const int grid = (int)(tiles < kSMs ? tiles : kSMs);
  HFL_LAUNCH((k_gather_gemm<<<grid, G_THREADS, G_SMEM, st>>>(tmap, p)));
  return HFL_OK;
}

}  // extern "C"
