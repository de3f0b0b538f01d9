// Microbenchmark: how fast can every SM stream an L2-resident weight matrix through a
// TMA ring of S stages?  (Round-1 question: is the fused MLP bound by L2->SM latency,
// i.e. bytes in flight, or by L2 bandwidth?)  Optionally 2-CTA clusters with multicast.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mb_arrive_cluster(uint32_t b, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 r;\n\tmapa.shared::cluster.u32 r, %0, %1;\n\t"
               "mbarrier.arrive.shared::cluster.b64 _, [r];\n\t}" ::"r"(b), "r"(cta) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
  uint32_t d = 0, spins = 0;
  while (!d) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(d) : "r"(b), "r"(ph) : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma2d(uint32_t dst, const void* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_mc(uint32_t dst, const void* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

template <int CL>
__global__ void __launch_bounds__(256) k_stream(const __grid_constant__ CUtensorMap tm, int S, int loads, int rows_total, int kcols, int box_rows, long long* prof) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4 * 32];
  const uint32_t stage_bytes = 16384;
  const int pair = threadIdx.x / 64, NP = blockDim.x / 64;
  uint32_t full = su32(bars + pair * 32), empty = su32(bars + pair * 32 + 16);
  const uint32_t ring = su32(smem) + pair * S * stage_bytes;
  loads /= NP;
  uint32_t rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x % 64 == 0) {
    for (int s = 0; s < S; ++s) { mb_init(full + 8 * s, 1); mb_init(empty + 8 * s, CL); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const int nkb = kcols / 64, nrb = rows_total / 128;
  if (threadIdx.x % 64 == 0) {                            // producer
    long long t_w = 0, t_e = 0, t_t = 0;
    for (int g = 0; g < loads; ++g) {
      const int s = g % S, ph = (g / S) & 1;
      long long c0_ = clock64();
      mb_wait(empty + 8 * s, ph ^ 1);
      long long c1_ = clock64();
      mb_expect(full + 8 * s, stage_bytes);
      long long c2_ = clock64();
      const int box = (g * NP + pair + (CL > 1 ? blockIdx.x / CL : blockIdx.x) * 7) % (nkb * nrb);
      const int c0 = (box % nkb) * 64, c1 = (box / nkb) * 128;
      if (CL == 1) tma2d(ring + s * stage_bytes, &tm, full + 8 * s, c0, c1);
      else tma2d_mc(ring + s * stage_bytes + rank * (stage_bytes / CL), &tm, full + 8 * s, c0, c1 + rank * box_rows, (uint16_t)((1 << CL) - 1));
      long long c3_ = clock64();
      t_w += c1_ - c0_; t_e += c2_ - c1_; t_t += c3_ - c2_;
    }
    if (blockIdx.x == 0 && pair == 0 && prof) { prof[0] = t_w; prof[1] = t_e; prof[2] = t_t; }
  } else if (threadIdx.x % 64 == 32) {                    // consumer
    long long t_cw = 0, t_ca = 0;
    for (int g = 0; g < loads; ++g) {
      const int s = g % S, ph = (g / S) & 1;
      long long c0_ = clock64();
      mb_wait(full + 8 * s, ph);
      long long c1_ = clock64();
      if (CL == 1) mb_arrive(empty + 8 * s);
      else for (int c = 0; c < CL; ++c) mb_arrive_cluster(empty + 8 * s, c);
      long long c2_ = clock64();
      t_cw += c1_ - c0_; t_ca += c2_ - c1_;
    }
    if (blockIdx.x == 0 && pair == 0 && prof) { prof[3] = t_cw; prof[4] = t_ca; }
  }
  __syncthreads();
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static long long* g_prof;
template <int CL>
static float run(const CUtensorMap& tm, int S, int loads, int rows, int kcols, int grid, int NP) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 * NP); cfg.dynamicSmemBytes = NP * S * 16384;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(k_stream<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 16384));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, k_stream<CL>, tm, S, loads, rows, kcols, 128 / CL, g_prof));
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) CK(cudaLaunchKernelEx(&cfg, k_stream<CL>, tm, S, loads, rows, kcols, 128 / CL, g_prof));
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const int rows = 2048, kcols = 512;                      // 2 MB bf16: W1 + W2 of one C=256 MLP
  void* w; CK(cudaMalloc(&w, (size_t)rows * kcols * 2)); CK(cudaMemset(w, 0, (size_t)rows * kcols * 2));
  EncodeFn enc; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int loads = 4096;
  CK(cudaMalloc(&g_prof, 64)); CK(cudaMemset(g_prof, 0, 64));
  for (int CL = 1; CL <= 2; ++CL) {
    CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(128 / CL)}; cuuint32_t el[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int grid : {8, 148}) for (int NP : {1, 2, 4}) for (int S : {1, 3}) {
      float ms = CL == 1 ? run<1>(tm, S, loads, rows, kcols, grid, NP) : run<2>(tm, S, loads, rows, kcols, grid, NP);
      long long hp[5]; CK(cudaMemcpy(hp, g_prof, 40, cudaMemcpyDeviceToHost));
      const double L = loads / NP;
      printf("   cycles/load: prod wait %.0f expect %.0f tma %.0f | cons wait %.0f arrive %.0f\n", hp[0] / L, hp[1] / L, hp[2] / L, hp[3] / L, hp[4] / L);
      double per_sm = (double)loads * 16384 / (ms * 1e-3) / 1e9;
      printf("cluster=%d grid=%3d pairs=%d stages=%2d  %.3f ms  %.1f GB/s per SM  %.2f TB/s aggregate (SM ingress)\n", CL, grid, NP, S, ms, per_sm, per_sm * grid / 1e3);
    }
  }
  return 0;
}
