// Microbenchmark 6: issue rate of small tcgen05.mma instructions when consecutive MMAs target
// DIFFERENT accumulators (round-robin over nacc tiles) vs the same one -- can independent small MMAs
// pipeline?  M = 128, K = 16, N = 16 / 64 / 128, SS and TS forms, unrolled issue loop.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../hotformerloc_b200/csrc/ptx.cuh"
using namespace hfl;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NACC, int CHAIN, int TS>
__global__ void __launch_bounds__(160) k_rr(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t sA = base, sB = base + 16384;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint64_t ad = ptx::umma_desc_sw128(sA), bd = ptx::umma_desc_sw128(sB);
    const uint32_t stride = 512 / (NACC + 1) / 16 * 16 < (uint32_t)N ? (uint32_t)N : 32;   // accumulator spacing (columns)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // CHAIN consecutive MMAs per accumulator, then the next accumulator
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
#pragma unroll
        for (int k = 0; k < CHAIN; ++k) {
          if (TS) ptx::umma_bf16_ts(tm + a * stride, tm + 448 + 8 * (k & 3), bd + 2 * (k & 3), idesc, 1);
          else ptx::umma_bf16(tm + a * stride, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
        }
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

template <int NACC, int CHAIN, int TS>
void run(int N, long long* out) {
  const int iters = 1024;
  CK(cudaFuncSetAttribute(k_rr<NACC, CHAIN, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int r = 0; r < 2; ++r) { k_rr<NACC, CHAIN, TS><<<148, 160, 65536>>>(N, iters, out); CK(cudaDeviceSynchronize()); }
  long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
  printf("N=%3d %s accumulators=%d chain=%d: %.1f cycles/MMA (ideal %.1f)\n", N, TS ? "TS" : "SS", NACC, CHAIN,
         (double)cyc / ((double)iters * NACC * CHAIN), 128.0 * N * 16 * 2 / 8192.0);
}

int main() {
  long long* out; CK(cudaMalloc(&out, 64));
  for (int N : {16, 64, 128}) {
    if (N <= 64) {
      run<1, 8, 0>(N, out); run<2, 1, 0>(N, out); run<4, 1, 0>(N, out); run<6, 1, 0>(N, out);
      run<2, 4, 0>(N, out); run<4, 2, 0>(N, out); run<4, 4, 0>(N, out); run<4, 8, 0>(N, out);
      run<1, 8, 1>(N, out); run<4, 1, 1>(N, out); run<4, 8, 1>(N, out);
    } else {
      run<1, 8, 0>(N, out); run<2, 1, 0>(N, out); run<3, 1, 0>(N, out); run<2, 4, 0>(N, out);
    }
  }
  return 0;
}
