// Fused MLP half-block for the narrow Swin stages (C = 128, 256):
//     x[M, C] += GELU( xn[M, C] * W1[4C, C]^T + b1 ) * W2[C, 4C]^T + b2          (x fp32 in place, xn / W 16-bit)
// Replaces intermediate.dense + GELU + output.dense + residual add (HF:swin/modeling_swin.py:510-531, 650) and,
// above all, the [M, 4C] hidden tensor: at stage 0 / batch 256 that is 822 MB written and read back per block,
// which made fc1 and fc2 the two most expensive kernels of the stage.  Here the hidden activations exist only as
// 128 x 128 chunks: fp32 in TMEM after GEMM1, GELU'd to 16 bit in registers, staged in shared memory in the
// tcgen05 K-major layout, and consumed as the A operand of GEMM2.
//
// One CTA owns a 128-token tile and walks the 4C/128 hidden chunks:
//   warp 0     TMA: the xn tile once per tile, then W1 / W2 as 128x64 boxes in the order the MMAs need them
//   warp 1     MMA issuer: G1(j+1) is issued before G2(j), so the tensor core works on the next chunk while the
//              epilogue warps apply GELU to the current one
//   warps 2-9  chunk epilogue (TMEM -> +b1 -> GELU -> 16 bit -> smem) and, after the last chunk, the fp32
//              residual epilogue of the tile (same coalesced path as the GEMM engine)
// TMEM: two 128-column GEMM1 accumulators + one C-column GEMM2 accumulator (<= 512 columns).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"

namespace csvit {

constexpr uint32_t kUnit = 128 * 128;   // bytes of one 128-row x 64-column 16-bit box (A slab / weight unit)

template <int C>
struct MlpCfg {
  static constexpr int KB1 = C / 64;            // k-blocks of GEMM1
  static constexpr int N2 = C / 128;            // 128-column groups of the GEMM2 output
  static constexpr int NCH = 4 * C / 128;       // hidden chunks
  static constexpr int WSTAGES = 4;             // (C = 128 gave two of its six weight stages to the second staging buffer)
  static constexpr int STG_BUFS = C == 128 ? 2 : 1;   // C = 128: residual in / result out by TMA (tma_f32 epilogue), prefetched at tile start
  static constexpr uint32_t A1_BYTES = KB1 * kUnit;
  static constexpr uint32_t A2_BYTES = 2 * 2 * kUnit;
  static constexpr uint32_t W_BYTES = WSTAGES * kUnit;
  static constexpr uint32_t STG_BYTES = STG_BUFS * kEpiWarps * kStageBufBytes;
  static constexpr size_t SMEM = 1024 + size_t(A1_BYTES) + A2_BYTES + W_BYTES + STG_BYTES + 512;
};

struct MlpParams {
  const float* b1;
  int M;
  long long* trace;      // debugging (CSVIT_MLP_TRACE): clock64 stamps of CTA 0's first tiles, else null
};
constexpr int kMlpTraceTiles = 8;
#ifdef CSVIT_MLP_TRACE_BUILD      // make EXTRA=-DCSVIT_MLP_TRACE_BUILD: clock64 stamps (they cost registers and a few per cent)
#define MLP_STAMP(k) do { if (tr) tr[k] = clock64(); } while (0)
#else
#define MLP_STAMP(k) do { } while (0)
#endif

// SPLIT (C = 128): the chunk epilogue (GELU) and the tile's residual epilogue run on DIFFERENT warps and the GEMM2 accumulator is
// double-buffered in TMEM (2 x 128 + 2 x 128 columns), so tile t's residual epilogue (~3500 cycles, mostly TMA waits) and the wait for
// its last GEMM2 (~1100) overlap the GELU chunks of tile t + 1 instead of preceding them on the same eight warps
// (profiles/r2_mlp_fused_trace.txt: 4 x 2000 + 1100 + 3500 cycles per tile in series).
// Warp roles with SPLIT (warpgroup-aligned for setmaxnreg): 0 TMA, 1 MMA, 2-3 idle, 4-11 chunk epilogue, 12-19 residual epilogue.
constexpr int kMlpThreadsSplit = 640;
template <int FMT, int C, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kMlpThreadsSplit : kGemmThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmR, MlpParams mp, EpiParams ep) {
  using Cfg = MlpCfg<C>;
  static_assert(!SPLIT || C == 128, "the second GEMM2 accumulator fits TMEM at C = 128 only");
  constexpr int KB1 = Cfg::KB1, N2 = Cfg::N2, NCH = Cfg::NCH, WS = Cfg::WSTAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a1_off = 0, a2_off = Cfg::A1_BYTES, w_off = a2_off + Cfg::A2_BYTES, stg_off = w_off + Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stg_off + Cfg::STG_BYTES);
  uint64_t* w_full = bars;                    // [WS]
  uint64_t* w_empty = bars + WS;              // [WS]
  uint64_t* acc1_full = bars + 2 * WS;        // [2]
  uint64_t* acc1_empty = acc1_full + 2;       // [2]
  uint64_t* a2_full = acc1_full + 4;          // [2]
  uint64_t* a2_empty = acc1_full + 6;         // [2]
  uint64_t* a1_full = acc1_full + 8;
  uint64_t* a1_empty = acc1_full + 9;
  uint64_t* acc2_full = acc1_full + 10;       // [2] (SPLIT: one per GEMM2 accumulator; else only [0])
  uint64_t* acc2_empty = acc1_full + 12;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 14);
  uint64_t* rbar = bars + 32;                 // [kEpiWarps][2] residual-chunk arrivals (tma_f32 epilogue, C = 128)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (mp.M + kBM - 1) / kBM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    if (ep.tma_f32) prefetch_tmap(&tmR);
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&rbar[i], 1);
    for (int s = 0; s < WS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1); mbar_init(&acc1_empty[b], kEpiWarps);
      mbar_init(&a2_full[b], kEpiWarps); mbar_init(&a2_empty[b], 1);
    }
    mbar_init(a1_full, 1); mbar_init(a1_empty, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc2_full[b], 1); mbar_init(&acc2_empty[b], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();      // PDL: the next kernel's prologue may overlap this kernel's tail ...
  griddep_wait();        // ... and this kernel touches global memory only after its predecessors have completed
  const uint32_t tm_acc2 = tmem_base + 256;
  constexpr int EW0 = SPLIT ? 4 : 2;         // first chunk-epilogue warp
  // SPLIT: the register pool is the launch allocation (640 x 96): warps 0-3 hand back 128 x 32 and the residual-epilogue warps 256 x 16 for
  // 256 x 32 more in the GELU warps; every setmaxnreg sits at the top of its role's branch so that ptxas allocates that branch against the new limit
  if (warp < EW0) {
  if (SPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");      // (40 made the issuer spill)
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      auto load_w = [&](const CUtensorMap* tm, int col, int row) {
        mbar_wait(&w_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&w_full[s], kUnit);
        tma_load_2d(smem + w_off + size_t(s) * kUnit, tm, &w_full[s], col, row);
        if (++s == WS) { s = 0; ph ^= 1u; }
      };
      uint32_t it = 0;
      auto load_a1 = [&](int t, uint32_t i) {       // the xn tile of the CTA's i-th tile (single buffer: free once the previous tile's last GEMM1 completes)
        mbar_wait(a1_empty, (i & 1u) ^ 1u);
        mbar_arrive_expect_tx(a1_full, Cfg::A1_BYTES);
        for (int kb = 0; kb < KB1; ++kb) tma_load_2d(smem + a1_off + size_t(kb) * kUnit, &tmX, a1_full, kb * 64, t * kBM);
      };
      bool pre = false;      // the next tile's xn and GEMM1(0) weights were requested ahead (same order as the issuer's, below)
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        if (!pre) {
          load_a1(t, it);
          for (int kb = 0; kb < KB1; ++kb) load_w(&tmW1, kb * 64, 0);                     // G1(0)
        }
        pre = false;
        for (int j = 0; j < NCH; ++j) {
          if (j + 1 < NCH) {
            for (int kb = 0; kb < KB1; ++kb) load_w(&tmW1, kb * 64, (j + 1) * 128);       // G1(j+1)
          } else if (SPLIT && t + int(gridDim.x) < num_tiles) {
            load_a1(t + int(gridDim.x), it + 1);                                          // next tile: xn, then G1(0)
            for (int kb = 0; kb < KB1; ++kb) load_w(&tmW1, kb * 64, 0);
            pre = true;
          }
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int n2 = 0; n2 < N2; ++n2) load_w(&tmW2, j * 128 + kb2 * 64, n2 * 128);   // G2(j)
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (converged warp, one elected lane issues: see fa_elect_one) ----------------
    {
      constexpr uint32_t idesc = make_idesc(uint32_t(FMT), kBM, 128);
      int s = 0; uint32_t ph = 0;
      auto mma_unit = [&](uint32_t a_addr, uint32_t d_tmem, bool first) {
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        if (fa_elect_one()) {
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(base + w_off + uint32_t(s) * kUnit);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss<false>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (first && k == 0) ? 0u : 1u);
          umma_commit(&w_empty[s]);
        }
        __syncwarp();
        if (++s == WS) { s = 0; ph ^= 1u; }
      };
      uint32_t it = 0;
      uint32_t use[2] = {0, 0};   // how many chunks have used acc1 / A2 buffer b so far (all tiles)
      auto gemm1 = [&](int j) {
        const int b = j & 1;
        mbar_wait(&acc1_empty[b], (use[b] & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB1; ++kb) mma_unit(base + a1_off + uint32_t(kb) * kUnit, tmem_base + uint32_t(b * 128), kb == 0);
        if (fa_elect_one()) umma_commit(&acc1_full[b]);
        __syncwarp();
      };
      bool pre = false;      // GEMM1(0) of this tile was issued during the previous tile's last chunk
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        if (!pre) {
          mbar_wait(a1_full, it & 1u);
          tc_fence_after();
          gemm1(0);
        }
        pre = false;
        for (int j = 0; j < NCH; ++j) {
          const int b = j & 1;
          if (j + 1 < NCH) {
            gemm1(j + 1);
            if (j + 2 == NCH) { if (fa_elect_one()) umma_commit(a1_empty); __syncwarp(); }   // last GEMM1 of the tile issued: xn tile free once it completes
          } else if (SPLIT && t + int(gridDim.x) < num_tiles) {
            // "G1(j+1) before G2(j)" across the tile boundary (SPLIT only: with one set of epilogue warps the next tile's weight units
            // would only delay this tile's last GEMM2 and its residual epilogue): the next tile's first GEMM1 goes ahead of this tile's last GEMM2, which
            // waits for the last GELU chunk - otherwise the GELU warps idle ~1500 cycles at every tile start (clock64 trace)
            mbar_wait(a1_full, (it + 1) & 1u);
            tc_fence_after();
            gemm1(0);
            pre = true;
          }
          mbar_wait(&a2_full[b], use[b] & 1u);
          tc_fence_after();
          const uint32_t ab = SPLIT ? (it & 1u) : 0u;             // GEMM2 accumulator of this tile
          if (j == 0) { mbar_wait(&acc2_empty[ab], ((SPLIT ? it >> 1 : it) & 1u) ^ 1u); tc_fence_after(); }
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int n2 = 0; n2 < N2; ++n2)
              mma_unit(base + a2_off + uint32_t(b * 2 + kb2) * kUnit, tm_acc2 + ab * 128u + uint32_t(n2 * 128), j == 0 && kb2 == 0);
          if (fa_elect_one()) {
            umma_commit(&a2_empty[b]);
            if (j == NCH - 1) umma_commit(&acc2_full[ab]);
          }
          __syncwarp();
          ++use[b];
        }
      }
    }
  }
  } else if (SPLIT && warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // ---------------- residual epilogue of tile `it` from GEMM2 accumulator it & 1, while the GELU warps work on tile it + 1 ----------------
    const int e = warp - 12;
    const int quad = warp & 3;
    const int half = e >> 2;
    uint8_t* stg = smem + stg_off + e * Cfg::STG_BUFS * kStageBufBytes;
    uint32_t it = 0, rph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      // the tma_f32 form of gemm.cuh::epilogue_tile for this kernel's one case (C = 128: two 32 x 32 fp32 chunks per warp, both
      // residual chunks requested at tile start) - written out here so that this branch compiles against 96 registers without spills
      const uint32_t ab = it & 1u;
      const int row0 = t * kBM + quad * 32, colbase = half * 64;
      if (ep.red_add) {      // the result leaves by TMA reduce-add: no residual chunks to request, only the staging buffers to get back
        if (lane == 0) tma_store_wait_read0();
        __syncwarp();
      } else {
        tma_f32_prefetch<C>(ep, &tmR, stg, rbar + 2 * e, t, 0, quad, half, lane);
      }
      mbar_wait(&acc2_full[ab], (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tm_acc2 + ab * 128u + (uint32_t(quad * 32) << 16) + uint32_t(colbase);
#pragma unroll 1
      for (int ci = 0; ci < 2; ++ci) {
        uint8_t* buf = stg + ci * kStageBufBytes;
        uint32_t rr[32];
        tmem_ld_32x32(t_addr + uint32_t(32 * ci), rr);
        tmem_ld_wait();
        if (row0 < ep.M && !ep.red_add) {
          mbar_wait(&rbar[2 * e + ci], (rph >> ci) & 1u);
          rph ^= (1u << ci);
        }
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + colbase + 32 * ci);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4* pp = reinterpret_cast<float4*>(buf + lane * 128 + ((k ^ (lane & 7)) << 4));
          const float4 bb = __ldg(b4 + k);
          float4 x0 = ep.red_add ? make_float4(0.f, 0.f, 0.f, 0.f) : *pp;
          x0.x += __uint_as_float(rr[4 * k]) + bb.x; x0.y += __uint_as_float(rr[4 * k + 1]) + bb.y;
          x0.z += __uint_as_float(rr[4 * k + 2]) + bb.z; x0.w += __uint_as_float(rr[4 * k + 3]) + bb.w;
          *pp = x0;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < ep.M) {
          if (ep.red_add) tma_reduce_add_2d(&tmR, buf, colbase + 32 * ci, row0);
          else tma_store_2d(&tmR, buf, colbase + 32 * ci, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[ab]);
    }
  } else {
    if (SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // ---------------- epilogues ----------------
    const int e = warp - EW0;
    const int quad = warp & 3;
    const int half = e >> 2;
    const int r = quad * 32 + lane;           // row of the tile owned by this thread
    const bool bf = FMT == 1;
    uint8_t* stg = smem + stg_off + e * Cfg::STG_BUFS * kStageBufBytes;
    uint32_t it = 0, rph = 0;
    uint32_t use[2] = {0, 0};
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      // residual chunks of this tile requested now: they land while the hidden chunks are being processed
      if (!SPLIT) {
        if (ep.tma_f32) tma_f32_prefetch<C>(ep, &tmR, stg, rbar + 2 * e, t, 0, quad, half, lane);
        else prefetch_resid_tile<C>(ep, t, 0, quad, half, lane);   // C = 256 (no room for a second staging buffer): at least pull the lines into L2
      }
#ifdef CSVIT_MLP_TRACE_BUILD
      long long* tr = (mp.trace && blockIdx.x == 0 && e == 0 && lane == 0 && it < kMlpTraceTiles) ? mp.trace + it * 64 : nullptr;
#endif
      for (int j = 0; j < NCH; ++j) {
        const int b = j & 1;
        MLP_STAMP(4 * j);
        mbar_wait(&acc1_full[b], use[b] & 1u);
        tc_fence_after();
        MLP_STAMP(4 * j + 1);
        uint32_t rr0[32], rr1[32];             // both 32-column TMEM loads are in flight before the single wait
        {
          const uint32_t ta = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(b * 128 + half * 64);
          tmem_ld_32x32(ta, rr0);
          tmem_ld_32x32(ta + 32u, rr1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);            // the accumulator is in registers: GEMM1(j+2) may overwrite it
        uint32_t pk[32];                      // this thread's 64 hidden columns, packed to 16 bit
        epi_bias_gelu_pack32(mp.b1, bf, j * 128 + half * 64, rr0, pk);
        epi_bias_gelu_pack32(mp.b1, bf, j * 128 + half * 64 + 32, rr1, pk + 16);
        MLP_STAMP(4 * j + 2);
        mbar_wait(&a2_empty[b], (use[b] & 1u) ^ 1u);           // GEMM2(j-2) has finished reading this A2 buffer
        MLP_STAMP(4 * j + 3);
        uint8_t* dst = smem + a2_off + uint32_t(b * 2 + half) * kUnit + r * 128;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
          *reinterpret_cast<uint4*>(dst + ((ci ^ (r & 7)) << 4)) = make_uint4(pk[4 * ci], pk[4 * ci + 1], pk[4 * ci + 2], pk[4 * ci + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[b]);
        ++use[b];
      }
      MLP_STAMP(4 * NCH);
      if (!SPLIT) {
#ifdef CSVIT_MLP_TRACE_BUILD
        if (tr) { mbar_wait(&acc2_full[0], it & 1u); tr[4 * NCH + 1] = clock64(); }
#endif
        // tile epilogue: acc2 (+ b2) + residual -> x, coalesced fp32 path of the GEMM engine (waits on acc2_full itself)
        epilogue_tile<C>(ep, &tmR, stg, tm_acc2, &acc2_full[0], it & 1u, t, 0, quad, half, lane, 1, nullptr, &tmR, rbar + 2 * e, &rph, true);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[0]);
      } else {
        MLP_STAMP(4 * NCH + 1);
      }
      MLP_STAMP(4 * NCH + 2);
    }
  }

  if (ep.tma_f32 && warp >= (SPLIT ? 12 : 2) && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int FMT, int C, bool SPLIT>
static int launch_mlp_s(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmR,
                      const MlpParams& mp, const EpiParams& ep, cudaStream_t stream) {
  using Cfg = MlpCfg<C>;
  static DeviceOnce once;
  auto kern = mlp_fused_kernel<FMT, C, SPLIT>;
  if (once.first()) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
  }
  const int tiles = (mp.M + kBM - 1) / kBM;
  static const int cap = [] { const char* e = getenv("CSVIT_MLP_CTAS"); return e ? atoi(e) : 0; }();      // ablation: fewer CTAs than SMs
  int ctas = tiles < num_sms() ? tiles : num_sms();
  if (cap > 0 && cap < ctas) ctas = cap;
  CSVIT_CUDA(launch_pdl(kern, dim3(ctas), dim3(SPLIT ? kMlpThreadsSplit : kGemmThreads), Cfg::SMEM, stream, tmX, tmW1, tmW2, tmR, mp, ep));
  return 0;
}
template <int FMT, int C>
static int launch_mlp(const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmR,
                      const MlpParams& mp, const EpiParams& ep, cudaStream_t stream) {
  if constexpr (C == 128) {      // CSVIT_MLP_SPLIT=0: chunk and residual epilogue on the same eight warps (ablation)
    static const bool split = [] { const char* e = getenv("CSVIT_MLP_SPLIT"); return !(e && e[0] == '0'); }();
    if (split && ep.tma_f32 && ep.resid && ep.bias) return launch_mlp_s<FMT, C, true>(tmX, tmW1, tmW2, tmR, mp, ep, stream);
  }
  return launch_mlp_s<FMT, C, false>(tmX, tmW1, tmW2, tmR, mp, ep, stream);
}

int launch_mlp_fused(const void* xn, long long ldxn, const void* W1, long long ldw1, const float* b1, const void* W2,
                     long long ldw2, const float* b2, float* x, long long ldx, int dtype, int M, int C, cudaStream_t stream) {
  CSVIT_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "mlp_fused: 16-bit operand formats only");
  CSVIT_REQUIRE(C == 128 || C == 256, "mlp_fused: C=%d not in {128, 256}", C);
  CSVIT_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "mlp_fused: residual stream must be 16-byte aligned");
  if (M <= 0) return 0;
  EpiParams ep{};
  ep.bias = b2; ep.resid = x; ep.out = x; ep.ldo = ldx; ep.ldr = ldx; ep.out_dtype = DT_F32; ep.act = ACT_NONE;
  ep.M = M; ep.N = C; ep.vec_ok = 1; ep.tma_store = 0; ep.coalesced = 1;
  ep.map_mode = ROWMAP_IDENTITY; ep.geom = make_geom(1, 1, 1, 0);
  MlpParams mp{b1, M, nullptr};
#ifdef CSVIT_MLP_TRACE_BUILD
  static const char* trace_path = getenv("CSVIT_MLP_TRACE");
#else
  static const char* trace_path = nullptr;      // the stamps are compiled out: nothing to record
#endif
  if (trace_path) { CSVIT_CUDA(cudaMalloc(&mp.trace, kMlpTraceTiles * 64 * sizeof(long long))); CSVIT_CUDA(cudaMemsetAsync(mp.trace, 0, kMlpTraceTiles * 64 * sizeof(long long), stream)); }
  CUtensorMap tmX, tmW1, tmW2;
  if (int e = make_tmap(&tmX, xn, ldxn, M, C, dtype, kBM, true)) return e;
  if (int e = make_tmap(&tmW1, W1, ldw1, 4ll * C, C, dtype, 128, true)) return e;
  if (int e = make_tmap(&tmW2, W2, ldw2, C, 4ll * C, dtype, 128, true)) return e;
  CUtensorMap tmR = tmX;
  static const bool tma_resid = [] { const char* e = getenv("CSVIT_MLP_TMA_RESID"); return !(e && e[0] == '0'); }();
  if (C == 128 && tma_resid && (ldx * 4) % 16 == 0) {
    // residual tile in / out by TMA: 32 x 32 fp32 boxes of x, both chunks of a warp requested at tile start
    ep.tma_f32 = 1; ep.coalesced = 0;
    if (int e = make_tmap(&tmR, x, ldx, M, C, DT_F32, 32, false)) return e;
  }
  // x is updated in place: the GEMM2 tile leaves as a reduction (gemm.cu, CSVIT_RED_ADD).  The split kernel keeps ep.resid as its dispatch
  // condition and looks at ep.red_add; the shared epilogue_tile() paths take resid == nullptr.
  static const bool split128 = [] { const char* e = getenv("CSVIT_MLP_SPLIT"); return !(e && e[0] == '0'); }();
  if (ep.tma_f32 ? red_add_mode() != 0 : red_add_mode() == 1) {
    ep.red_add = 1;
    if (!(C == 128 && ep.tma_f32 && split128)) ep.resid = nullptr;
  }
  const bool bf = dtype == DT_BF16;
  if (trace_path) {      // debugging: one traced launch, stamps written as text (cycles relative to the first stamp)
    int e = C == 128 ? (bf ? launch_mlp<1, 128>(tmX, tmW1, tmW2, tmR, mp, ep, stream) : launch_mlp<0, 128>(tmX, tmW1, tmW2, tmR, mp, ep, stream))
                     : (bf ? launch_mlp<1, 256>(tmX, tmW1, tmW2, tmR, mp, ep, stream) : launch_mlp<0, 256>(tmX, tmW1, tmW2, tmR, mp, ep, stream));
    if (e) return e;
    CSVIT_CUDA(cudaStreamSynchronize(stream));
    long long h[kMlpTraceTiles * 64];
    CSVIT_CUDA(cudaMemcpy(h, mp.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(mp.trace);
    if (FILE* f = fopen(trace_path, "a")) {
      const int nch = 4 * C / 128;
      fprintf(f, "# mlp_fused C=%d: epilogue warp 0 of CTA 0; per chunk: wait acc1 | TMEM load + GELU | wait A2 buffer | store; then wait acc2 | residual epilogue\n", C);
      for (int t = 0; t < kMlpTraceTiles; ++t) {
        const long long* r = h + t * 64;
        if (!r[0]) continue;
        fprintf(f, "tile %d start %8lld:", t, r[0] - h[0]);
        for (int j = 0; j < nch; ++j) {
          const long long nxt = j + 1 < nch ? r[4 * j + 4] : r[4 * nch];
          fprintf(f, "  [%lld %lld %lld %lld]", r[4 * j + 1] - r[4 * j], r[4 * j + 2] - r[4 * j + 1], r[4 * j + 3] - r[4 * j + 2], nxt - r[4 * j + 3]);
        }
        fprintf(f, "  acc2 wait %lld, epilogue %lld, total %lld\n", r[4 * nch + 1] - r[4 * nch], r[4 * nch + 2] - r[4 * nch + 1], r[4 * nch + 2] - r[0]);
      }
      fclose(f);
    }
    return 0;
  }
  if (C == 128) return bf ? launch_mlp<1, 128>(tmX, tmW1, tmW2, tmR, mp, ep, stream) : launch_mlp<0, 128>(tmX, tmW1, tmW2, tmR, mp, ep, stream);
  return bf ? launch_mlp<1, 256>(tmX, tmW1, tmW2, tmR, mp, ep, stream) : launch_mlp<0, 256>(tmX, tmW1, tmW2, tmR, mp, ep, stream);
}

}  // namespace csvit
