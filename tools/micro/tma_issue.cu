// Microbenchmark 2: one shared ring of S 16 KB stages, consumed by ONE thread; the TMA loads
// are issued by P threads that take stages round-robin, either lanes of one warp (mode 0) or
// lane 0 of P different warps (mode 1).  Question: what serialises back-to-back TMA loads?
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
__global__ void __launch_bounds__(320) k_issue(const __grid_constant__ CUtensorMap tm, int S, int loads, int P, int mode, int box_rows, int nkb, int nrb, int prefetch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[64];
  const uint32_t stage_bytes = 64 * 2 * box_rows;
  uint32_t full = su32(bars), empty = su32(bars + 32);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mb_init(full + 8 * s, 1); mb_init(empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int me = -1;                                            // issuer id
  if (mode == 0) { if (warp == 1 && lane < P) me = lane; }
  else { if (warp >= 1 && warp <= P && lane == 0) me = warp - 1; }
  if (me >= 0) {
    if (prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    for (int g = me; g < loads; g += P) {
      const int s = g % S, ph = (g / S) & 1;
      mb_wait(empty + 8 * s, ph ^ 1);
      mb_expect(full + 8 * s, stage_bytes);
      const int box = (g + blockIdx.x * 7) % (nkb * nrb);
      tma2d(su32(smem) + s * stage_bytes, &tm, full + 8 * s, (box % nkb) * 64, (box / nkb) * box_rows);
    }
  } else if (threadIdx.x == 0) {
    for (int g = 0; g < loads; ++g) {
      const int s = g % S, ph = (g / S) & 1;
      mb_wait(full + 8 * s, ph);
      mb_arrive(empty + 8 * s);
    }
  }
  __syncthreads();
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int rows = 2048, kcols = 512;
  void* w; CK(cudaMalloc(&w, (size_t)rows * kcols * 2)); CK(cudaMemset(w, 0, (size_t)rows * kcols * 2));
  EncodeFn enc; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  CK(cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384));
  for (int box_rows : {64, 128, 256}) {
    CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t el[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    const int stage = 128 * box_rows, S = 12 * 16384 / stage > 12 ? 12 : 12 * 16384 / stage;
    const int loads = 2048 * 16384 / stage * 2;
    for (int mode : {0, 1}) for (int P : {1, 2, 4, 8}) for (int pf : {1}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 2; ++i) k_issue<<<148, 320, S * stage>>>(tm, S, loads, P, mode, box_rows, kcols / 64, rows / box_rows, pf);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) k_issue<<<148, 320, S * stage>>>(tm, S, loads, P, mode, box_rows, kcols / 64, rows / box_rows, pf);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("box 64x%-3d (%2d KB) stages=%2d %s issuers=%d: %.3f ms  %.1f GB/s per SM\n", box_rows, stage / 1024, S, mode ? "warps" : "lanes", P, ms, (double)loads * stage / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
