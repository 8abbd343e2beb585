// Microbenchmark: tcgen05.ld (LDTM) throughput per SM with 4 / 8 / 16 warps reading 32-lane x 32-column fp32 blocks.
// The fused window-attention kernel drains ~190 TMEM columns per (tile, head); this tells whether that is a bound.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cs-vit_b200/csrc tools/cuda/tmem_ld_rate.cu -o tools/bin/tmem_ld_rate
#include <cstdio>
#include "common.cuh"
using namespace csvit;

__global__ void __launch_bounds__(512, 1) ld_rate(int iters, long long* cycles, unsigned* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t addr = tm + (uint32_t((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(addr + uint32_t((it * 64) & 448), r0);
    tmem_ld_32x32(addr + uint32_t((it * 64 + 32) & 480), r1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r0[j] + r1[j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; unsigned* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int grid : {1, 148}) {
      ld_rate<<<grid, warps * 32, 0>>>(iters, d, s);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * 2 * 32 * 32 * 4 * warps;
      printf("warps=%2d grid=%3d: %s  %.1f B/clk/SM  (%.1f cycles per 4 KB warp-load)\n", warps, grid, cudaGetErrorString(e),
             bytes / double(c), double(c) / (iters * 2));
    }
  }
  return 0;
}
