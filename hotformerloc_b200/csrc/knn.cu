// Exact L2 top-k retrieval: replaces the faiss
// IndexFlatL2 / sklearn KDTree search of eval/pnv_evaluate.py:200-225.  fp32
// squared differences on the CUDA cores (recall parity needs fp32-accurate ranking;
// the work is ~N_q * N_db * 256 FMAs, i.e. sub-millisecond at evaluation sizes).
// A rank searches its shard of the database (rows [idx_offset, idx_offset + n_db)) and
// returns per-query partial top-k lists sorted by (distance, index); hfl_topk_merge
// folds the all-gathered partial lists into the global top-k.  Inside a rank the database is
// split once more over the grid (hfl_knn_topk_ws) so that small query sets still fill the GPU.
#include "common.cuh"

namespace hfl {

constexpr int KN_TQ = 64, KN_TD = 64, KN_KC = 16, KN_MAXK = 32, KN_LD = 68, KN_MAXSPLIT = 16;

// grid = (query tiles of 64, database splits): every CTA ranks its 64 queries against its slice of the
// database and writes a sorted partial top-k list [split][nq][k]; k_topk_merge folds the splits (and, in the
// sharded evaluation, the ranks).  Per 64 x 64 distance tile the 256 threads hold 4 x 4 fp32 squared
// distances each; a candidate survives only if it is not worse than the query's current k-th best (kept in
// shared memory), survivors are compacted per query with a shared-memory counter and inserted by the query's
// own thread -- after the first few tiles almost nothing survives, so the tile loop is pure FMA work instead of
// 64 threads walking every candidate while 192 wait.  Ordering is the total order (distance, index): the
// result does not depend on the order in which survivors were compacted.
__global__ void __launch_bounds__(256)
k_knn(const float* __restrict__ q, int nq, const float* __restrict__ db, int ndb, int dim, int k,
      int idx_offset, int per_split, float* __restrict__ out_d, int32_t* __restrict__ out_i) {
  __shared__ __align__(16) float sq[KN_KC][KN_LD];
  __shared__ __align__(16) float sd[KN_KC][KN_LD];
  __shared__ float cand_d[KN_TQ][KN_TD];
  __shared__ uint8_t cand_j[KN_TQ][KN_TD];
  __shared__ int cnt[KN_TQ];
  __shared__ float worst_d[KN_TQ];
  __shared__ float best_d[KN_TQ][KN_MAXK];
  __shared__ int32_t best_i[KN_TQ][KN_MAXK];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * KN_TQ;
  const int lo = blockIdx.y * per_split, hi = min(ndb, lo + per_split);
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  for (int i = tid; i < KN_TQ * KN_MAXK; i += 256) {
    best_d[i / KN_MAXK][i % KN_MAXK] = INFINITY;
    best_i[i / KN_MAXK][i % KN_MAXK] = 0x7fffffff;
  }
  if (tid < KN_TQ) { cnt[tid] = 0; worst_d[tid] = INFINITY; }
  for (int d0 = lo; d0 < hi; d0 += KN_TD) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int c0 = 0; c0 < dim; c0 += KN_KC) {
      __syncthreads();
      // 64 rows x 16 columns of the query and the database tile, transposed into [column][row]; a warp reads
      // 16 consecutive columns of a row (64 B) per half-warp
      for (int i = tid; i < KN_TQ * KN_KC; i += 256) {
        const int r = i / KN_KC, c = i % KN_KC;
        sq[c][r] = (q0 + r < nq && c0 + c < dim) ? __ldg(q + (size_t)(q0 + r) * dim + c0 + c) : 0.f;
        sd[c][r] = (d0 + r < hi && c0 + c < dim) ? __ldg(db + (size_t)(d0 + r) * dim + c0 + c) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < KN_KC; ++c) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sq[c][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&sd[c][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) { const float e = a[u] - b[v]; acc[u][v] = fmaf(e, e, acc[u][v]); }
      }
    }
    // survivors of the tile, compacted per query (<= also admits index ties; the insertion decides them)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = ty * 4 + u;
      const float w = worst_d[r];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int col = tx * 4 + v;
        if (acc[u][v] <= w && d0 + col < hi && q0 + r < nq) {
          const int pos = atomicAdd(&cnt[r], 1);
          cand_d[r][pos] = acc[u][v];
          cand_j[r][pos] = (uint8_t)col;
        }
      }
    }
    __syncthreads();
    if (tid < KN_TQ && cnt[tid] > 0) {
      float* bd = best_d[tid];
      int32_t* bi = best_i[tid];
      const int n = cnt[tid];
      for (int j = 0; j < n; ++j) {
        const float dj = cand_d[tid][j];
        const int32_t ij = idx_offset + d0 + (int)cand_j[tid][j];
        if (dj < bd[k - 1] || (dj == bd[k - 1] && ij < bi[k - 1])) {
          int pos = k - 1;
          while (pos > 0 && (bd[pos - 1] > dj || (bd[pos - 1] == dj && bi[pos - 1] > ij))) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
          }
          bd[pos] = dj;
          bi[pos] = ij;
        }
      }
      cnt[tid] = 0;
      worst_d[tid] = bd[k - 1];
    }
    // (the __syncthreads at the top of the next tile's first chunk orders these updates before its filter)
  }
  __syncthreads();
  float* od = out_d + (size_t)blockIdx.y * nq * k;
  int32_t* oi = out_i + (size_t)blockIdx.y * nq * k;
  for (int i = tid; i < KN_TQ * k; i += 256) {
    const int r = i / k, j = i % k;
    if (q0 + r < nq) {
      od[(size_t)(q0 + r) * k + j] = best_d[r][j];
      oi[(size_t)(q0 + r) * k + j] = best_i[r][j] == 0x7fffffff ? -1 : best_i[r][j];
    }
  }
}

// merge `parts` sorted partial lists per query: in [parts, nq, k] -> out [nq, k]
__global__ void k_topk_merge(const float* __restrict__ in_d, const int32_t* __restrict__ in_i,
                             int parts, int nq, int k, float* __restrict__ out_d,
                             int32_t* __restrict__ out_i) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  int head[16];
  for (int p = 0; p < parts; ++p) head[p] = 0;
  for (int j = 0; j < k; ++j) {
    int bp = -1;
    float bd = INFINITY;
    int32_t bi = 0x7fffffff;
    for (int p = 0; p < parts; ++p) {
      if (head[p] >= k) continue;
      const size_t o = ((size_t)p * nq + qi) * k + head[p];
      const float d = in_d[o];
      const int32_t i = in_i[o];
      if (i < 0) continue;
      if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; bp = p; }
    }
    if (bp >= 0) ++head[bp];
    out_d[(size_t)qi * k + j] = bp >= 0 ? bd : INFINITY;
    out_i[(size_t)qi * k + j] = bp >= 0 ? bi : -1;
  }
}

}  // namespace hfl

using namespace hfl;

extern "C" {

// splits of the database such that the grid fills the GPU (>= 2 CTAs per SM) without slices below 8 tiles
static int knn_splits(int nq, int ndb) {
  const int qt = (nq + KN_TQ - 1) / KN_TQ;
  int s = (2 * sm_count() + qt - 1) / qt;
  const int max_by_db = (ndb + 8 * KN_TD - 1) / (8 * KN_TD);
  if (s > max_by_db) s = max_by_db;
  if (s > KN_MAXSPLIT) s = KN_MAXSPLIT;
  return s < 1 ? 1 : s;
}

int64_t hfl_knn_workspace_bytes(int32_t nq, int32_t ndb, int32_t k) {
  const int s = knn_splits(nq, ndb);
  return s > 1 ? (int64_t)s * nq * k * 8 : 0;
}

int hfl_knn_topk_ws(const float* q, int32_t nq, const float* db, int32_t ndb, int32_t dim, int32_t k,
                    int32_t idx_offset, float* out_d, int32_t* out_i, void* ws, int64_t ws_bytes, void* stream_) {
  if (nq == 0) return HFL_OK;
  HFL_CHECK_ARG(q && db && out_d && out_i, "null argument");
  HFL_CHECK_ARG(k >= 1 && k <= KN_MAXK, "k must be in [1, 32]");
  cudaStream_t st = (cudaStream_t)stream_;
  int s = knn_splits(nq, ndb);
  if (!ws || ws_bytes < (int64_t)s * nq * k * 8) s = 1;          // no workspace: one slice per query tile
  const int qt = (nq + KN_TQ - 1) / KN_TQ;
  int per = (ndb + s - 1) / s;
  per = (per + KN_TD - 1) / KN_TD * KN_TD;
  if (s == 1) {
    HFL_LAUNCH((k_knn<<<dim3(qt, 1), 256, 0, st>>>(q, nq, db, ndb, dim, k, idx_offset, per, out_d, out_i)));
    return HFL_OK;
  }
  float* pd = (float*)ws;
  int32_t* pi = (int32_t*)((char*)ws + (size_t)s * nq * k * 4);
  HFL_LAUNCH((k_knn<<<dim3(qt, s), 256, 0, st>>>(q, nq, db, ndb, dim, k, idx_offset, per, pd, pi)));
  HFL_LAUNCH((k_topk_merge<<<(nq + 127) / 128, 128, 0, st>>>(pd, pi, s, nq, k, out_d, out_i)));
  return HFL_OK;
}

int hfl_knn_topk(const float* q, int32_t nq, const float* db, int32_t ndb, int32_t dim, int32_t k,
                 int32_t idx_offset, float* out_d, int32_t* out_i, void* stream_) {
  return hfl_knn_topk_ws(q, nq, db, ndb, dim, k, idx_offset, out_d, out_i, nullptr, 0, stream_);
}

int hfl_topk_merge(const float* in_d, const int32_t* in_i, int32_t parts, int32_t nq, int32_t k,
                   float* out_d, int32_t* out_i, void* stream_) {
  if (nq == 0) return HFL_OK;
  HFL_CHECK_ARG(in_d && in_i && out_d && out_i, "null argument");
  HFL_CHECK_ARG(parts >= 1 && parts <= 16, "at most 16 partial lists");
  HFL_LAUNCH((k_topk_merge<<<(nq + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(in_d, in_i, parts, nq, k, out_d, out_i)));
  return HFL_OK;
}

}  // extern "C"
