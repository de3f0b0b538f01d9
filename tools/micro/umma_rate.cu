// Microbenchmark 3: tcgen05.mma issue/execute rate of ONE CTA per SM with operands resident in
// shared memory (no loads at all): N = 64/128/256, accumulating into one or two TMEM tiles,
// optionally with concurrent shared-memory traffic from the other warps (to emulate an
// epilogue / producers hammering smem).  cycles per MMA vs the 8192 FLOP/clk/SM peak.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../hotformerloc_b200/csrc/ptx.cuh"
using namespace hfl;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(288) k_umma(int N, int iters, int nacc, int traffic, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t sA = base, sB = base + 16384;              // A: 128 x 64 bf16 (one K block); B: up to 256 x 64
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
      const uint64_t ad = ptx::umma_desc_sw128(sA), bd = ptx::umma_desc_sw128(sB);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tm + (uint32_t)((it % nacc) * N);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1);
      }
      ptx::umma_commit(ptx::smem_u32(&bar));
      ptx::mbar_wait(ptx::smem_u32(&bar), 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  } else if (traffic) {
    // the other 8 warps stream 16-byte smem stores+loads over a private 32 KB region
    uint32_t a = base + 65536 + (uint32_t)(warp - 1) * 4096 + lane * 16;
    uint32_t acc = 0;
    volatile uint64_t* b = &bar;
    for (int r = 0; r < traffic; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a + i * 512), "r"(acc) : "memory");
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a + i * 512) : "memory");
        acc += x + y + z + w;
      }
    }
    if (acc == 0x12345678u) out[1] = acc + (long long)*b;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out; CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768));
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 4096;
  for (int grid : {1, 148}) for (int N : {64, 128, 256}) for (int nacc : {1, 2}) for (int traffic : {0, 20000}) {
    if (nacc * N > 512) continue;
    k_umma<<<grid, 288, 65536 + 32768>>>(N, iters, nacc, traffic, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_umma<<<grid, 288, 65536 + 32768>>>(N, iters, nacc, traffic, out);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; CK(cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost));
    const double per = (double)cyc / (iters * 4), ideal = 128.0 * N * 16 * 2 / 8192.0;
    printf("grid=%3d N=%3d acc=%d smem-traffic=%d: %.1f cycles/MMA (ideal %.0f) -> %.0f%% of peak;  kernel %.3f ms\n", grid, N, nacc, traffic ? 1 : 0, per, ideal, 100 * ideal / per, ms);
  }
  return 0;
}
