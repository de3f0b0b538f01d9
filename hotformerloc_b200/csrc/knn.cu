// Exact L2 top-k retrieval: replaces the faiss
// IndexFlatL2 / sklearn KDTree search of eval/pnv_evaluate.py:200-225.  fp32
// squared differences on the CUDA cores (recall parity needs fp32-accurate ranking;
// the work is ~N_q * N_db * 256 FMAs, i.e. sub-millisecond at evaluation sizes).
// A rank searches its shard of the database (rows [idx_offset, idx_offset + n_db)) and
// returns per-query partial top-k lists sorted by (distance, index); hfl_topk_merge
// folds the all-gathered partial lists into the global top-k.
#include "common.cuh"

namespace hfl {

constexpr int KN_TQ = 64, KN_TD = 64, KN_KC = 16, KN_MAXK = 32;

__global__ void __launch_bounds__(256)
k_knn(const float* __restrict__ q, int nq, const float* __restrict__ db, int ndb, int dim, int k,
      int idx_offset, float* __restrict__ out_d, int32_t* __restrict__ out_i) {
  __shared__ float sq[KN_KC][KN_TQ + 1];
  __shared__ float sd[KN_KC][KN_TD + 1];
  __shared__ float dist[KN_TQ][KN_TD + 1];
  __shared__ float best_d[KN_TQ][KN_MAXK];
  __shared__ int32_t best_i[KN_TQ][KN_MAXK];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * KN_TQ;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
  for (int i = tid; i < KN_TQ * KN_MAXK; i += 256) {
    best_d[i / KN_MAXK][i % KN_MAXK] = INFINITY;
    best_i[i / KN_MAXK][i % KN_MAXK] = 0x7fffffff;
  }
  for (int d0 = 0; d0 < ndb; d0 += KN_TD) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int c0 = 0; c0 < dim; c0 += KN_KC) {
      __syncthreads();
      for (int i = tid; i < KN_TQ * KN_KC; i += 256) {
        const int r = i / KN_KC, c = i % KN_KC;
        sq[c][r] = (q0 + r < nq && c0 + c < dim) ? q[(size_t)(q0 + r) * dim + c0 + c] : 0.f;
        sd[c][r] = (d0 + r < ndb && c0 + c < dim) ? db[(size_t)(d0 + r) * dim + c0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < KN_KC; ++c) {
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = sq[c][ty * 4 + u]; b[u] = sd[c][tx * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) { const float e = a[u] - b[v]; acc[u][v] += e * e; }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) dist[ty * 4 + u][tx * 4 + v] = acc[u][v];
    __syncthreads();
    if (tid < KN_TQ && q0 + tid < nq) {
      float* bd = best_d[tid];
      int32_t* bi = best_i[tid];
      const int lim = min(KN_TD, ndb - d0);
      for (int j = 0; j < lim; ++j) {
        const float dj = dist[tid][j];
        const int32_t ij = idx_offset + d0 + j;
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
    }
  }
  __syncthreads();
  for (int i = tid; i < KN_TQ * k; i += 256) {
    const int r = i / k, j = i % k;
    if (q0 + r < nq) {
      out_d[(size_t)(q0 + r) * k + j] = best_d[r][j];
      out_i[(size_t)(q0 + r) * k + j] = best_i[r][j] == 0x7fffffff ? -1 : best_i[r][j];
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

int hfl_knn_topk(const float* q, int32_t nq, const float* db, int32_t ndb, int32_t dim, int32_t k,
                 int32_t idx_offset, float* out_d, int32_t* out_i, void* stream_) {
  if (nq == 0) return HFL_OK;
  HFL_CHECK_ARG(q && db && out_d && out_i, "null argument");
  HFL_CHECK_ARG(k >= 1 && k <= KN_MAXK, "k must be in [1, 32]");
  HFL_LAUNCH((k_knn<<<(nq + KN_TQ - 1) / KN_TQ, 256, 0, (cudaStream_t)stream_>>>(q, nq, db, ndb, dim, k, idx_offset, out_d, out_i)));
  return HFL_OK;
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
