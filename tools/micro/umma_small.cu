// Microbenchmark 5: small-N tcgen05.mma rates and the M = 64 accumulator layout, for the design of the
// tensor-core attention core (head_dim 16: S = Q K^T is one K = 16 MMA, P V has N = 16).
//   (a) cycles per MMA for M = 128 / 64, N = 16 .. 128, SS and TS (A operand from tensor memory) forms
//   (b) which TMEM lanes receive the 64 rows of an M = 64 MMA (printed as a lane -> row map)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../hotformerloc_b200/csrc/ptx.cuh"
using namespace hfl;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(160) k_rate(int M, int N, int ts, int iters, int ksteps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t sA = base, sB = base + 16384;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, N);
    const uint64_t ad = ptx::umma_desc_sw128(sA), bd = ptx::umma_desc_sw128(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + (uint32_t)((it & 1) * 128);
      for (int k = 0; k < ksteps; ++k) {
        if (ts) ptx::umma_bf16_ts(d, tm + 256 + 8 * (k & 7), bd + 2 * (k & 3), idesc, 1);
        else ptx::umma_bf16(d, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
      }
    }
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

// M = 64 layout probe: A[r][0] = r + 1, B[n][0] = 1  =>  D[r][n] = r + 1
__global__ void __launch_bounds__(128) k_layout(int M, float* out /*[128 lanes][2]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t sA = base, sB = base + 16384;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int r = threadIdx.x;
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + ((0 ^ (r & 7)) << 4));
    a[0] = __float2bfloat16((float)(r + 1));
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(smem + 16384 + r * 128 + ((0 ^ (r & 7)) << 4));
    b[0] = __float2bfloat16(1.0f);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tptr;
  // clear 32 columns of every lane
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-7.0f);
    ptx::tmem_st32(tm + ((uint32_t)(warp * 32) << 16), z);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, 16);
    ptx::umma_bf16(tm, ptx::umma_desc_sw128(sA), ptx::umma_desc_sw128(sB), idesc, 0);
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t r[32];
  ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  out[threadIdx.x * 2] = __uint_as_float(r[0]);
  out[threadIdx.x * 2 + 1] = __uint_as_float(r[15]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out; CK(cudaMalloc(&out, 64));
  float* lay; CK(cudaMalloc(&lay, 128 * 2 * 4));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int iters = 2048;
  for (int M : {128, 64}) for (int N : {16, 32, 64, 128}) for (int ts : {0, 1}) for (int ks : {1, 8}) {
    if (ts && M == 64) continue;
    k_rate<<<148, 160, 65536>>>(M, N, ts, iters, ks, out);
    CK(cudaDeviceSynchronize());
    k_rate<<<148, 160, 65536>>>(M, N, ts, iters, ks, out);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    printf("M=%3d N=%3d %s k-steps/accumulator=%d: %.1f cycles/MMA (ideal %.1f)\n", M, N, ts ? "TS" : "SS", ks,
           (double)cyc / ((double)iters * ks), (double)M * N * 16 * 2 / 8192.0);
  }
  for (int M : {128, 64}) {
    k_layout<<<1, 128, 65536>>>(M, lay);
    CK(cudaDeviceSynchronize());
    float h[256]; CK(cudaMemcpy(h, lay, sizeof(h), cudaMemcpyDeviceToHost));
    printf("M=%d accumulator layout, lane: row+1 (col 0 | col 15), -7 = untouched\n", M);
    for (int l = 0; l < 128; ++l) printf("%d:%g|%g%s", l, h[2 * l], h[2 * l + 1], (l % 16 == 15) ? "\n" : "  ");
  }
  return 0;
}
