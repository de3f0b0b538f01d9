// Micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (bf16 pairs packed into 32-bit
// columns, lane = row) and B in shared memory (K-major, SWIZZLE_128B).  Verifies the operand
// layout the fused MLP relies on for GEMM 2 (H kept in TMEM).  D = A[128x64] * B[128x64]^T.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../hotformerloc_b200/csrc/ptx.cuh"
using namespace hfl;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_ts(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int a_col0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t sB = ptx::smem_u32(smem);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, r = threadIdx.x;
  // B: row n = threadIdx.x, 8 chunks of 16 B
  for (int c = 0; c < 8; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    *reinterpret_cast<uint4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tptr;
  // A row r -> TMEM lane r, packed columns a_col0 .. a_col0+31
  uint32_t w[32];
  for (int k = 0; k < 32; ++k) w[k] = *reinterpret_cast<const uint32_t*>(A + r * 64 + 2 * k);
  ptx::tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + a_col0, w);
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
    const uint64_t bd = ptx::umma_desc_sw128(sB);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tm + 256, tm + a_col0 + 8 * k, bd + 2 * k, idesc, k != 0);
    ptx::umma_commit(ptx::smem_u32(&bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t raw[32];
    ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + 256 + c0, raw);
    ptx::tmem_ld_wait();
    for (int k = 0; k < 32; ++k) D[r * 128 + c0 + k] = __uint_as_float(raw[k]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  __nv_bfloat16 *hA = new __nv_bfloat16[128 * 64], *hB = new __nv_bfloat16[128 * 64];
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, 128 * 64 * 2)); CK(cudaMalloc(&dB, 128 * 64 * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaMemcpy(dA, hA, 128 * 64 * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, 128 * 64 * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  for (int a_col0 : {0, 64}) {
    k_ts<<<1, 128, 16384>>>(dA, dB, dD, a_col0);
    CK(cudaDeviceSynchronize());
    float* hD = new float[128 * 128];
    CK(cudaMemcpy(hD, dD, 128 * 128 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)__bfloat162float(hA[m * 64 + k]) * (double)__bfloat162float(hB[n * 64 + k]);
      maxerr = fmax(maxerr, fabs(s - hD[m * 128 + n]));
    }
    printf("A in TMEM at column %d: max |D - ref| = %.3e  (%s)\n", a_col0, maxerr, maxerr < 1e-3 ? "layout OK" : "MISMATCH");
  }
  return 0;
}
