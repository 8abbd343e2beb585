// LayerNorm-prologue GEMM on CTA pairs, A-stationary:
//     out[M, N] = act( LN(x)[M, C] * W[N, C]^T + bias )            (16-bit out, TMA-stored)
// for the two Linears that follow a LayerNorm in every Swin block: Q/K/V (rows gathered in shifted-window
// order) and MLP fc1 (+GELU).  Replaces layernorm_before + pad + roll + window_partition + 3 addmm
// (HF:swin/modeling_swin.py:606-622, 404-406) and layernorm_after + addmm + gelu (HF:648, 514-519) without ever
// writing the normalised activations to HBM.
//
// Each CTA of a pair normalises ITS 128 rows once per 256-row block - fp32 rows read straight from the residual
// stream (two-pass statistics in registers, same arithmetic as ln_rows_kernel) and written as 16-bit into
// shared memory in the 128-byte-swizzled K-major layout that tcgen05 reads - and keeps that A tile resident
// while the whole width N streams past it: only the weight tile (half per CTA, cta_group::2) enters the SM per
// MMA, 16 KB per 512 tensor cycles instead of 48 KB in the plain GEMM.
//
// Roles per CTA: warp 0 = TMA producer for W, warp 1 = MMA issuer (leader CTA only), warps 2-9 = LayerNorm
// producers for the block, then epilogue warps for its N/256 output tiles.
#include <type_traits>

#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"

namespace csvit {

constexpr int kLnBN = 256;
constexpr int kLnStages = 4;
constexpr uint32_t kLnBBytes = (kLnBN / 2) * 128;    // this CTA's half of a 256 x 64 weight slab
constexpr uint32_t kLnSlabBytes = kBM * 128;         // one 128-row x 64-column slab of A

template <int KB>
struct LnCfg {
  static constexpr uint32_t A_BYTES = KB * kLnSlabBytes;
  static constexpr uint32_t B_BYTES = kLnStages * kLnBBytes;
  static constexpr uint32_t STG_BYTES = kEpiWarps * kStageBufBytes;
  static constexpr size_t SMEM = 1024 + size_t(A_BYTES) + B_BYTES + STG_BYTES + 256;
};

struct LnParams {
  const float* x;       // fp32 residual stream [*, C]
  const float* gamma;
  const float* beta;
  float eps;
  int mode;             // LN_IDENTITY or LN_WINDOW
  WinGeom geom;
};

// FMT: 0 = fp16, 1 = bf16.  KB = C / 64.
template <int FMT, int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
lngemm_pair_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, LnParams ln, EpiParams ep) {
  using Cfg = LnCfg<KB>;
  using T16 = typename std::conditional<FMT == 1, __nv_bfloat16, __half>::type;
  constexpr int BN = kLnBN, STAGES = kLnStages, C = KB * 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a_tile = smem;
  uint8_t* b_tiles = smem + Cfg::A_BYTES;
  uint8_t* staging = b_tiles + Cfg::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STG_BYTES);
  uint64_t* bfull = bars;                     // [STAGES]  leader
  uint64_t* bempty = bars + STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;        // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]       leader
  uint64_t* afull = bars + 2 * STAGES + 4;    // [1]       leader: both CTAs' A tiles are written
  uint64_t* aempty = bars + 2 * STAGES + 5;   // [1]       all MMAs of the block have read A
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_mp = (ep.M + 2 * kBM - 1) / (2 * kBM);
  const int num_n = (ep.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bfull[s], 2); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * kEpiWarps); }
    mbar_init(afull, 2);
    mbar_init(aempty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- weight producer ----------------
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int mp = pair_id; mp < num_mp; mp += num_pairs) {
        for (int n_blk = 0; n_blk < num_n; ++n_blk) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bempty[s], ph ^ 1u);
            const uint32_t lfull = mapa_u32(smem_u32(&bfull[s]), 0);
            mbar_arrive_expect_tx_cluster(lfull, kLnBBytes);
            tma_load_2d_pair(b_tiles + size_t(s) * kLnBBytes, &tmB, lfull, kb * 64, n_blk * BN + int(rank) * (BN / 2));
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else {
      // Idle lanes: pull the NEXT block's residual-stream rows into L2 while this block is multiplied, paced by
      // the per-block `aempty` barrier so the prefetch runs exactly one block ahead of the LayerNorm warps.
      uint32_t pph = 0;
      int j = 0;
      for (int mp = pair_id; mp < num_mp; mp += num_pairs, ++j) {
        if (j > 0) { mbar_wait(aempty, pph); pph ^= 1u; }
        const int nmp = mp + num_pairs;
        if (nmp >= num_mp) break;
        const int m_blk = nmp * 2 + int(rank);
        for (int r = lane - 1; r < kBM; r += 31) {
          const int row = m_blk * kBM + r;
          if (row >= ep.M) continue;
          long long src = row;
          if (ln.mode == LN_WINDOW) {
            const int b = row / ln.geom.N, w = row - b * ln.geom.N;
            src = static_cast<long long>(b) * ln.geom.N + win_row_to_token(ln.geom, w);
          }
          const char* p = reinterpret_cast<const char*>(ln.x + src * C);
#pragma unroll
          for (int o = 0; o < C * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader) ----------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(uint32_t(FMT), 2 * kBM, BN);
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      uint32_t blk_ph = 0;
      for (int mp = pair_id; mp < num_mp; mp += num_pairs) {
        mbar_wait(afull, blk_ph);
        tc_fence_after();
        for (int n_blk = 0; n_blk < num_n; ++n_blk) {
          mbar_wait(&tempty[as], aph ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&bfull[s], ph);
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(base + uint32_t(kb) * kLnSlabBytes);
            const uint64_t bdesc = make_sw128_kmajor_desc(base + Cfg::A_BYTES + uint32_t(s) * kLnBBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss_pair(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(&bempty[s], 3);
            if (kb == KB - 1) umma_commit_pair(&tfull[as], 3);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
          if (++as == 2) { as = 0; aph ^= 1u; }
        }
        umma_commit_pair(aempty, 3);   // A may be overwritten once every MMA of this block has completed
        blk_ph ^= 1u;
      }
    }
  } else {
    // ---------------- LayerNorm producers, then epilogue ----------------
    const int e = warp - 2;
    const int quad = warp & 3;
    const int half = e >> 2;
    uint8_t* stg = staging + e * kStageBufBytes;
    int as = 0; uint32_t aph = 0;
    uint32_t blk_ph = 0;
    constexpr int CHUNKS = KB * 8;                 // 16-byte (8-column) chunks per row
    constexpr int CPL = (CHUNKS + 31) / 32;        // chunks per lane
    for (int mp = pair_id; mp < num_mp; mp += num_pairs) {
      const int m_blk = mp * 2 + int(rank);
      // ---- LayerNorm of rows [16e, 16e+16) of this CTA's 128-row block: two rows per step, the next step's
      //      loads are issued before the current step's reductions (register double buffer) ----
      float v[2][2][CPL][8];
      bool ok[2][2];
      auto load_rows = [&](int buf, int rr) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = m_blk * kBM + 16 * e + rr + u;
          ok[buf][u] = row < ep.M;
          long long src = row;
          if (ok[buf][u] && ln.mode == LN_WINDOW) {
            const int b = row / ln.geom.N, w = row - b * ln.geom.N;
            src = static_cast<long long>(b) * ln.geom.N + win_row_to_token(ln.geom, w);
          }
          const float* xr = ln.x + src * C;
#pragma unroll
          for (int t = 0; t < CPL; ++t) {
            const int c = lane + 32 * t;
            if (ok[buf][u] && c < CHUNKS) {
              const float4 p0 = *reinterpret_cast<const float4*>(xr + c * 8);
              const float4 p1 = *reinterpret_cast<const float4*>(xr + c * 8 + 4);
              v[buf][u][t][0] = p0.x; v[buf][u][t][1] = p0.y; v[buf][u][t][2] = p0.z; v[buf][u][t][3] = p0.w;
              v[buf][u][t][4] = p1.x; v[buf][u][t][5] = p1.y; v[buf][u][t][6] = p1.z; v[buf][u][t][7] = p1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[buf][u][t][i] = 0.f;
            }
          }
        }
      };
      auto norm_rows = [&](int buf, int rr) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = 16 * e + rr + u;
          float sum = 0.f;
#pragma unroll
          for (int t = 0; t < CPL; ++t)
            sum += ((v[buf][u][t][0] + v[buf][u][t][1]) + (v[buf][u][t][2] + v[buf][u][t][3])) +
                   ((v[buf][u][t][4] + v[buf][u][t][5]) + (v[buf][u][t][6] + v[buf][u][t][7]));
          const float mean = warp_sum(sum) * (1.0f / float(C));
          float sq = 0.f;
#pragma unroll
          for (int t = 0; t < CPL; ++t) {
            if (lane + 32 * t < CHUNKS) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { const float d = v[buf][u][t][i] - mean; sq = fmaf(d, d, sq); }
            }
          }
          const float rstd = rsqrtf(warp_sum(sq) * (1.0f / float(C)) + ln.eps);
#pragma unroll
          for (int t = 0; t < CPL; ++t) {
            const int c = lane + 32 * t;
            if (c < CHUNKS) {
              uint4 pk = make_uint4(0u, 0u, 0u, 0u);
              if (ok[buf][u]) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(ln.gamma + c * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(ln.gamma + c * 8 + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ln.beta + c * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ln.beta + c * 8 + 4));
                const float* w = v[buf][u][t];
                pk.x = Half16<T16>::pack((w[0] - mean) * rstd * g0.x + b0.x, (w[1] - mean) * rstd * g0.y + b0.y);
                pk.y = Half16<T16>::pack((w[2] - mean) * rstd * g0.z + b0.z, (w[3] - mean) * rstd * g0.w + b0.w);
                pk.z = Half16<T16>::pack((w[4] - mean) * rstd * g1.x + b1.x, (w[5] - mean) * rstd * g1.y + b1.y);
                pk.w = Half16<T16>::pack((w[6] - mean) * rstd * g1.z + b1.z, (w[7] - mean) * rstd * g1.w + b1.w);
              }
              // chunk c of row r -> slab c/8, 16-byte position (c%8) ^ (r%8) inside the 128-byte row
              *reinterpret_cast<uint4*>(a_tile + (c >> 3) * kLnSlabBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = pk;
            }
          }
        }
      };
      load_rows(0, 0);                               // raw rows do not depend on A being free: request them first
      mbar_wait(aempty, blk_ph ^ 1u);                // previous block's MMAs are done with A (passes at once on the first)
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (it + 1 < 8) load_rows((it + 1) & 1, 2 * (it + 1));
        norm_rows(it & 1, 2 * it);
      }
      fence_proxy_async_smem();                      // generic-proxy writes -> visible to the tensor core's async proxy
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 producer warps of this CTA
      if (e == 0 && lane == 0) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");   // once per block: this CTA's A tile is visible cluster-wide
        mbar_arrive_cluster(mapa_u32(smem_u32(afull), 0));
      }
      // ---- epilogue over the block's output tiles ----
      for (int n_blk = 0; n_blk < num_n; ++n_blk) {
        epilogue_tile<BN>(ep, &tmC, stg, tmem_base + uint32_t(as * BN), &tfull[as], aph, m_blk, n_blk, quad, half, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
      blk_ph ^= 1u;
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

template <int FMT, int KB>
static int launch_lng(const CUtensorMap& tmB, const CUtensorMap& tmC, const LnParams& ln, const EpiParams& ep, cudaStream_t stream) {
  using Cfg = LnCfg<KB>;
  static bool configured = false;
  auto kern = lngemm_pair_kernel<FMT, KB>;
  if (!configured) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
    configured = true;
  }
  const int num_mp = (ep.M + 2 * kBM - 1) / (2 * kBM);
  int pairs = num_sms() / 2;
  if (pairs > num_mp) pairs = num_mp;
  kern<<<pairs * 2, kGemmThreads, Cfg::SMEM, stream>>>(tmB, tmC, ln, ep);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_ln_gemm(const float* x, const float* gamma, const float* beta, float eps, int mode, const WinGeom& g,
                   const void* W, long long ldw, int dtype, int M, int N, int C, const float* bias, int act, void* out,
                   long long ldo, cudaStream_t stream) {
  CSVIT_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "ln_linear: 16-bit operand formats only");
  CSVIT_REQUIRE(C == 128 || C == 256 || C == 512, "ln_linear: C=%d not in {128,256,512}", C);
  CSVIT_REQUIRE(N % 64 == 0 && ldo % 8 == 0, "ln_linear: N=%d must be a multiple of 64 and ldo of 8", N);
  CSVIT_REQUIRE(mode == LN_IDENTITY || mode == LN_WINDOW, "ln_linear: bad gather mode %d", mode);
  if (M <= 0) return 0;
  EpiParams ep{};
  ep.bias = bias; ep.out = out; ep.ldo = ldo; ep.out_dtype = dtype; ep.act = act;
  ep.M = M; ep.N = N; ep.vec_ok = 1; ep.tma_store = 1; ep.coalesced = 0;
  ep.map_mode = ROWMAP_IDENTITY; ep.geom = make_geom(1, 1, 1, 0);
  LnParams ln{x, gamma, beta, eps, mode, g};
  CUtensorMap tmB, tmC;
  if (int e = make_tmap(&tmB, W, ldw, N, C, dtype, kLnBN / 2, true)) return e;
  if (int e = make_tmap(&tmC, out, ldo, M, N, dtype, 32, false)) return e;
  const bool bf = dtype == DT_BF16;
  if (C == 512) return bf ? launch_lng<1, 8>(tmB, tmC, ln, ep, stream) : launch_lng<0, 8>(tmB, tmC, ln, ep, stream);
  if (C == 256) return bf ? launch_lng<1, 4>(tmB, tmC, ln, ep, stream) : launch_lng<0, 4>(tmB, tmC, ln, ep, stream);
  return bf ? launch_lng<1, 2>(tmB, tmC, ln, ep, stream) : launch_lng<0, 2>(tmB, tmC, ln, ep, stream);
}

}  // namespace csvit
