// Microbenchmark: raw tcgen05.mma issue rate from resident shared memory (no TMA, no epilogue).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cs-vit_b200/csrc tools/cuda/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include "common.cuh"
using namespace csvit;

template <int BN, int KSTEPS>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long* cycles, int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t stage_bytes = 128 * 128 + BN * 128;
  for (int i = threadIdx.x; i < int(stages * stage_bytes / 4); i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(1, 128, BN);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = base + (it % stages) * stage_bytes;
      const uint64_t ad = make_sw128_kmajor_desc(sa), bd = make_sw128_kmajor_desc(sa + 128 * 128);
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) umma_ss<false>(tm + (it & 1) * BN % 512, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int BN>
void run(const char* name, int stages, int grid) {
  long long* d; cudaMalloc(&d, 8);
  size_t smem = 1024 + stages * (128 * 128 + BN * 128);
  cudaFuncSetAttribute(mma_rate<BN, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  mma_rate<BN, 4><<<grid, 128, smem>>>(iters, d, stages);
  cudaError_t e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("%s grid=%d stages=%d: %s  %.1f cycles per 128x%dx16 MMA (ideal %d)\n", name, grid, stages, cudaGetErrorString(e),
         double(c) / (iters * 4), BN, BN / 2);
}
int main() {
  run<256>("M128 N256", 1, 1);
  run<256>("M128 N256", 4, 1);
  run<256>("M128 N256", 4, 148);
  run<128>("M128 N128", 4, 148);
  run<96>("M128 N96", 4, 148);
  run<64>("M128 N64", 4, 148);
  run<32>("M128 N32", 4, 148);
  return 0;
}
