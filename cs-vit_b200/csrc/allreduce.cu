// Gradient allreduce over NVLink / NVSwitch peer memory for the data-parallel finetune step (BASELINE configs[3];
// ref:scripts/finetune.py:133-135 wraps the model in DistributedDataParallel, whose bucketed NCCL allreduce this replaces).
//
// The flat fp32 gradient buckets of cs_vit.train.GradReducer live in SYMMETRIC memory: every rank can address every peer's copy
// (torch.distributed._symmetric_memory does the IPC / multicast mapping; this kernel only receives plain pointers).  One launch
// per bucket on a side stream, while the backward kernels of the remaining layers keep the main stream busy:
//
//   barrier      every rank's bucket is complete (per-CTA flags in peer memory, release / acquire at system scope)
//   reduce       rank r owns shard r of the bucket.
//                NVSwitch multicast mapping available: multimem.ld_reduce.add.v4.f32 pulls the shard THROUGH the switch, which
//                adds the W copies in flight (NVLS), the owner scales by 1/W and multimem.st broadcasts the result to all W copies:
//                each GPU sends and receives bytes/W per phase instead of bytes (W-1)/W.
//                Otherwise (two-shot over P2P): the owner loads the shard from all W peers, sums in rank order, and stores the
//                averaged result into all W copies.  Only the owner ever touches shard r between the barriers, so no further sync.
//   barrier      all stores have landed everywhere: the next kernel on the stream may read the whole bucket.
//
// 256 threads and <= 40 registers per CTA with no shared memory, so the CTAs co-reside with the persistent 1-CTA-per-SM GEMM
// kernels of the backward pass instead of queueing behind them.  Summation order is fixed, and the result of a shard is computed
// once and broadcast: the replicas stay bit-identical.
#include "common.cuh"
#include "errors.h"
#include "rowops.cuh"

namespace csvit {

constexpr int AR_THREADS = 256;
constexpr int AR_MAX_WORLD = 8;
constexpr int AR_MAX_CTAS = 64;

struct ArParams {
  float* bufs[AR_MAX_WORLD];        // the bucket's address in every rank's symmetric buffer
  uint32_t* flags[AR_MAX_WORLD];    // [2][AR_MAX_CTAS][AR_MAX_WORLD] uint32 per rank, zero-initialised once
  float* mc;                        // multicast address of the bucket, or nullptr
  long long n;                      // floats (multiple of 4)
  int rank, world;
  float scale;
};

__device__ __forceinline__ void ar_put(uint32_t* addr) {      // 0 -> 1 on the peer's flag (it was consumed back to 0 by its last wait)
  uint32_t old;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void ar_wait(uint32_t* addr) {     // 1 -> 0 on my own flag
  uint32_t old;
  uint32_t spins = 0;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 1u && ++spins > (1u << 28)) __trap();      // a peer that never arrives must surface as an error, not a hang
  } while (old != 1u);
}
// CTA b of every rank meets CTA b of every other rank.
__device__ __forceinline__ void ar_barrier(const ArParams& p, int phase) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    const size_t slot = (size_t(phase) * AR_MAX_CTAS + blockIdx.x) * AR_MAX_WORLD;
    __threadfence_system();
    ar_put(p.flags[peer] + slot + p.rank);
    ar_wait(p.flags[p.rank] + slot + peer);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(AR_THREADS, 6)
allreduce_f32_kernel(ArParams p) {
  ar_barrier(p, 0);
  const long long n4 = p.n >> 2;
  const long long per = (n4 + p.world - 1) / p.world;
  const long long lo = per * p.rank, hi = (lo + per < n4) ? lo + per : n4;
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  if (p.mc != nullptr) {
    // four switch-side reductions in flight per thread: the achieved rate is bytes-in-flight / NVLS latency
    for (long long i = lo + static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x; i < hi; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < hi) v[u] = mc_ld_reduce(p.mc + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < hi) {
          v[u].x *= p.scale; v[u].y *= p.scale; v[u].z *= p.scale; v[u].w *= p.scale;
          mc_st(p.mc + 4 * (i + u * stride), v[u]);
        }
    }
  } else {
    for (long long i = lo + static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x; i < hi; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s0 = 0; s0 < AR_MAX_WORLD; s0 += 4) {      // four peers' loads in flight at a time (register budget), rank order kept
        float4 v[4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < p.world) v[s] = __ldcv(reinterpret_cast<const float4*>(p.bufs[s0 + s] + 4 * i));      // never from a stale L1 line
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s0 + s < p.world) { acc.x += v[s].x; acc.y += v[s].y; acc.z += v[s].z; acc.w += v[s].w; }
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
#pragma unroll
      for (int s = 0; s < AR_MAX_WORLD; ++s)
        if (s < p.world) *reinterpret_cast<float4*>(p.bufs[s] + 4 * i) = acc;
    }
  }
  ar_barrier(p, 1);
}

int launch_allreduce_f32(void* const* bufs, void* const* flags, void* mc, long long n, int rank, int world, float scale, int ctas,
                         cudaStream_t stream) {
  CSVIT_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce: rank %d / world %d (1..%d)", rank, world,
                AR_MAX_WORLD);
  CSVIT_REQUIRE(n >= 0 && (n & 3) == 0, "allreduce: element count %lld must be a multiple of 4", n);
  if (n == 0) return 0;
  ArParams p{};
  for (int s = 0; s < world; ++s) {
    CSVIT_REQUIRE(bufs[s] != nullptr && flags[s] != nullptr && (reinterpret_cast<uintptr_t>(bufs[s]) & 15) == 0,
                  "allreduce: peer %d buffer / flag pointer missing or not 16-byte aligned", s);
    p.bufs[s] = static_cast<float*>(bufs[s]);
    p.flags[s] = static_cast<uint32_t*>(flags[s]);
  }
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(mc) & 15) == 0, "allreduce: multicast pointer not 16-byte aligned");
  p.mc = static_cast<float*>(mc);
  p.n = n; p.rank = rank; p.world = world; p.scale = scale;
  if (ctas <= 0) ctas = 32;      // measured (2 x B200, graphed finetune step): 32 CTAs hide the whole reduction behind backward; 64 move a
                                 // lone bucket faster (374 vs 218 GB/s) but take SMs from the backward GEMMs (0.75 ms exposed)
  if (ctas > AR_MAX_CTAS) ctas = AR_MAX_CTAS;
  // every rank must launch the same grid (CTA b meets CTA b): the size depends only on arguments that are equal on all ranks
  allreduce_f32_kernel<<<ctas, AR_THREADS, 0, stream>>>(p);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace csvit
