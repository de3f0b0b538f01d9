// Attention cores: head_dim is 16 for every
// stage of every shipped configuration, so one score tile is a 16x8x16 MMA and the
// softmax / relative-position bias dominate -- these kernels keep everything of one
// (window, head) in registers + a warp-private smem slice and never materialise the
// (N_win,K,K) masks or the (N_win,K,K,3) rel-pos tensors of the reference
// (models/octree.py:193-209, 272-283): the batch mask and the RPE index are computed
// from a per-token (x,y,z,submap) int16x4 table.
//   k_window_attn : octree window attention, plain / dilated / hierarchical (+relay token)
//   k_varlen_attn : relay-token self-attention over ragged per-submap sequences
// qkv is the bf16 output of the tcgen05 projection GEMM, laid out [row, 3C] as
// [q | k | v] x [head, 16]  (octformer_backbone.py:71-72).
#include "common.cuh"
#include "ptx.cuh"

namespace hfl {

constexpr int AT_HD = 16;
constexpr int AT_NT = 10;              // score n-tiles (8 keys each) -> up to 80 keys per pass
constexpr int AT_KEYS = AT_NT * 8;
constexpr int AT_RS = 48;              // smem row stride in bytes (16 bf16 + pad, conflict-free)
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

constexpr int WA_WARPS = 16;
constexpr int WA_WARP_SMEM = 2 * AT_KEYS * AT_RS + AT_KEYS * 8;   // K, V, xyzb

__global__ void __launch_bounds__(WA_WARPS * 32, 1) k_window_attn(const WinAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int L = p.K + (p.hat ? 1 : 0);
  const int num = 2 * p.bnd + 1;
  const int tbl = 3 * num;
  // RPE table transposed to [H][tbl] in smem
  float* s_rpe = reinterpret_cast<float*>(smem);
  const int rpe_bytes = p.rpe ? ((p.H * tbl * 4 + 15) & ~15) : 0;
  if (p.rpe) {
    for (int i = threadIdx.x; i < p.H * tbl; i += blockDim.x) {
      int h = i / tbl, e = i - h * tbl;
      s_rpe[i] = __ldg(p.rpe + (size_t)e * p.H + h) * LOG2E;
    }
  }
  __syncthreads();
  uint8_t* wbase = smem + rpe_bytes + (size_t)warp * WA_WARP_SMEM;
  uint8_t* sK = wbase;
  uint8_t* sV = wbase + AT_KEYS * AT_RS;
  short4* sT = reinterpret_cast<short4*>(wbase + 2 * AT_KEYS * AT_RS);
  const uint32_t sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);
  const int C3 = 3 * p.C;
  const float sc = p.scale * LOG2E;
  const int64_t items = (int64_t)p.n_win * p.H;
  const int n_mt = (L + 15) / 16;

  for (int64_t item = (int64_t)blockIdx.x * WA_WARPS + warp; item < items;
       item += (int64_t)gridDim.x * WA_WARPS) {
    const int w = (int)(item / p.H), h = (int)(item % p.H);
    // ---- stage K, V (this head) and the token table of the window ----
    for (int s = lane; s < AT_KEYS; s += 32) {
      int64_t row = 0, tok = -1;
      const bool ok = s < L;
      if (ok) slot_row(p, w, s, row, tok);
      const __nv_bfloat16* src = p.qkv + row * C3 + p.C + h * AT_HD;
      ptx::cp_async16(sK_u + s * AT_RS, src, ok ? 16u : 0u);
      ptx::cp_async16(sK_u + s * AT_RS + 16, src + 8, ok ? 16u : 0u);
      ptx::cp_async16(sV_u + s * AT_RS, src + p.C, ok ? 16u : 0u);
      ptx::cp_async16(sV_u + s * AT_RS + 16, src + p.C + 8, ok ? 16u : 0u);
      short4 tk = make_short4(0, 0, 0, -2);
      if (ok) {
        if (tok >= 0) tk = p.xyzb[tok];
        else { tk = p.xyzb[(int64_t)w * p.K]; tk.x = tk.y = tk.z = 0; }   // RT: id of first token
      }
      sT[s] = tk;
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<0>();
    __syncwarp();
    const float* tab = s_rpe + h * tbl;

    for (int mt = 0; mt < n_mt; ++mt) {
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      // ---- Q fragments straight from global ----
      uint32_t qa[4] = {0u, 0u, 0u, 0u};
      int64_t row0 = 0, row1 = 0, tk_;
      if (i0 < L) {
        slot_row(p, w, i0, row0, tk_);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p.qkv + row0 * C3 + h * AT_HD);
        qa[0] = __ldg(q + t); qa[2] = __ldg(q + t + 4);
      }
      if (i1 < L) {
        slot_row(p, w, i1, row1, tk_);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p.qkv + row1 * C3 + h * AT_HD);
        qa[1] = __ldg(q + t); qa[3] = __ldg(q + t + 4);
      }
      const short4 ti0 = sT[min(i0, AT_KEYS - 1)], ti1 = sT[min(i1, AT_KEYS - 1)];
      const bool rt0 = p.hat && i0 == 0;      // relay-token row: no RPE
      // ---- S = Q K^T ----
      float s[AT_NT][4];
#pragma unroll
      for (int nt = 0; nt < AT_NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        uint32_t kb[2];
        const uint8_t* kr = sK + (nt * 8 + g) * AT_RS + t * 4;
        kb[0] = *reinterpret_cast<const uint32_t*>(kr);
        kb[1] = *reinterpret_cast<const uint32_t*>(kr + 16);
        ptx::mma16816(s[nt], qa, kb);
      }
      // ---- bias: submap mask + relative position ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < AT_NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          const short4 tj = sT[j];
          const bool rtj = p.hat && j == 0;
          float b0 = 0.f, b1 = 0.f;
          if (p.rpe) {
            if (!rtj) {
              if (!rt0) {
                int dx = min(max((int)ti0.x - (int)tj.x, -p.bnd), p.bnd) + p.bnd;
                int dy = min(max((int)ti0.y - (int)tj.y, -p.bnd), p.bnd) + p.bnd + num;
                int dz = min(max((int)ti0.z - (int)tj.z, -p.bnd), p.bnd) + p.bnd + 2 * num;
                b0 = tab[dx] + tab[dy] + tab[dz];
              }
              int dx = min(max((int)ti1.x - (int)tj.x, -p.bnd), p.bnd) + p.bnd;
              int dy = min(max((int)ti1.y - (int)tj.y, -p.bnd), p.bnd) + p.bnd + num;
              int dz = min(max((int)ti1.z - (int)tj.z, -p.bnd), p.bnd) + p.bnd + 2 * num;
              b1 = tab[dx] + tab[dy] + tab[dz];
            }
          }
          const bool v0 = (j < L) && (i0 < L) && (tj.w == ti0.w);
          const bool v1 = (j < L) && (i1 < L) && (tj.w == ti1.w);
          s[nt][e] = v0 ? s[nt][e] * sc + b0 : -INFINITY;
          s[nt][2 + e] = v1 ? s[nt][2 + e] * sc + b1 : -INFINITY;
          mx0 = fmaxf(mx0, s[nt][e]);
          mx1 = fmaxf(mx1, s[nt][2 + e]);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      if (mx0 == -INFINITY) mx0 = 0.f;
      if (mx1 == -INFINITY) mx1 = 0.f;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < AT_NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mx0); s[nt][1] = exp2f(s[nt][1] - mx0);
        s[nt][2] = exp2f(s[nt][2] - mx1); s[nt][3] = exp2f(s[nt][3] - mx1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // ---- O = P V ----
      float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int kt = 0; kt < AT_NT / 2; ++kt) {
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
        ptx::mma16816(o[0], pa, b0);
        ptx::mma16816(o[1], pa, b1);
      }
      const float r0 = l0 > 0.f ? 1.f / l0 : 0.f, r1 = l1 > 0.f ? 1.f / l1 : 0.f;
      if (i0 < L) {
        uint32_t* d = reinterpret_cast<uint32_t*>(p.out + row0 * p.C + h * AT_HD);
        d[t] = pack_bf16(o[0][0] * r0, o[0][1] * r0);
        d[t + 4] = pack_bf16(o[1][0] * r0, o[1][1] * r0);
      }
      if (i1 < L) {
        uint32_t* d = reinterpret_cast<uint32_t*>(p.out + row1 * p.C + h * AT_HD);
        d[t] = pack_bf16(o[0][2] * r1, o[0][3] * r1);
        d[t + 4] = pack_bf16(o[1][2] * r1, o[1][3] * r1);
      }
    }
    __syncwarp();
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
constexpr int VA_WARPS = 4;
constexpr int VA_MT = 4;        // m-tiles per warp -> 256 query rows per CTA

__global__ void __launch_bounds__(VA_WARPS * 32) k_varlen_attn(const VarAttnParams p) {
  __shared__ __align__(16) uint8_t sK[AT_KEYS * AT_RS];
  __shared__ __align__(16) uint8_t sV[AT_KEYS * AT_RS];
  __shared__ int32_t sId[AT_KEYS];
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

  for (int kb0 = 0; kb0 < L; kb0 += AT_KEYS) {
    __syncthreads();
    for (int s = threadIdx.x; s < AT_KEYS; s += blockDim.x) {
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
      float s[AT_NT][4];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < AT_NT; ++nt) {
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
      for (int nt = 0; nt < AT_NT; ++nt) {
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
      for (int kt = 0; kt < AT_NT / 2; ++kt) {
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
  HFL_CHECK_ARG(C == H * AT_HD, "head_dim must be 16");
  HFL_CHECK_ARG(K + (hat ? 1 : 0) <= AT_KEYS, "window (+relay token) must fit 80 keys");
  HFL_CHECK_ARG(dil >= 1 && (!hat || dil == 1), "dilation is not used with relay tokens");
  HFL_CHECK_ARG(n_win % dil == 0, "window count must be a multiple of the dilation");
  WinAttnParams p;
  p.qkv = (const __nv_bfloat16*)qkv; p.out = (__nv_bfloat16*)out; p.xyzb = (const short4*)xyzb;
  p.rpe = rpe; p.n_win = (int)n_win; p.H = H; p.C = C; p.K = K; p.dil = dil; p.hat = hat;
  p.bnd = bnd; p.scale = scale;
  const int tbl = 3 * (2 * bnd + 1);
  const int rpe_bytes = rpe ? ((H * tbl * 4 + 15) & ~15) : 0;
  const int smem = rpe_bytes + WA_WARPS * WA_WARP_SMEM;
  HFL_CHECK_ARG(smem <= 227 * 1024, "RPE table too large for shared memory");
  static int smem_set = 0;
  if (smem > smem_set) {
    HFL_CUDA(cudaFuncSetAttribute(k_window_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  const int64_t items = n_win * H;
  int grid = (int)ceil_div(items, WA_WARPS);
  if (grid > kSMs) grid = kSMs;
  HFL_LAUNCH((k_window_attn<<<grid, WA_WARPS * 32, smem, st>>>(p)));
  return HFL_OK;
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
