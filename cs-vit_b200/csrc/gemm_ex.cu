// General-layout GEMM for the training step: D[M,N] (+)= sum_k A(m,k) * B(n,k) where each operand may be stored
// K-major ([MN, K] rows, the nn.Linear layout) or MN-major ([K, MN] rows).  That covers the two backward GEMMs of
// every Linear without materialising a transpose:
//   dgrad  dX[T, in]   = dY[T, out] * W[out, in]      A = dY (K-major),  B = W   (MN-major: stored [k = out, n = in])
//   wgrad  dW[out, in] = dY[T, out]^T * X[T, in]      A = dY (MN-major), B = X   (MN-major), k = tokens
// (torch autograd runs these as `mm` on transposed views for HF:swin/modeling_swin.py:404-406,479,514,527 and
// ref:cs_vit/net/transformer_module.py:262-264,282,290-294.)
//
// Same warp-specialised pipeline as gemm.cu (TMA producer / tcgen05 issuer / 8 epilogue warps, 128x128 tiles, 2-deep
// TMEM ring).  MN-major operands use the canonical SWIZZLE_128B MN-major shared-memory layout
//   ((8 x 16 B, n), (8, k)) : ((1, LBO), (8, SBO))   [units of 16 B]
// i.e. one TMA box of {128 B of MN, BK k-rows} per 128-byte MN chunk: LBO = chunk stride = BK*128 B, SBO = 1024 B.
// wgrad has a tiny output and a very deep K (all tokens), so work items are (tile, k-split) and the split partials
// are reduced with fp32 `red.global.add` into a zeroed (or accumulating) output.
#include "errors.h"
#include "gemm.cuh"

namespace csvit {

constexpr int kExBN = 128;
constexpr int kExStages = 5;
constexpr uint32_t kExABytes = kBM * 128;
constexpr uint32_t kExBBytes = kExBN * 128;
constexpr uint32_t kExStageBytes = kExABytes + kExBBytes;
constexpr uint32_t kExTiles = kExStages * kExStageBytes;
constexpr uint32_t kExStg = 2 * kEpiWarps * kStageBufBytes;
constexpr size_t kExSmem = 1024 + size_t(kExTiles) + kExStg + 256;

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t chunk_stride_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_stride_bytes >> 4) << 16;   // LBO: next 128-byte chunk along M/N
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // SBO: next group of 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                          // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_major(uint32_t fmt, int M, int N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

template <int FMT, bool A_MN, bool B_MN>  // FMT: 0 = fp16, 1 = bf16, 2 = tf32
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ex_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int K, int splits,
               int atomic, EpiParams ep) {
  constexpr bool TF32 = FMT == 2;
  constexpr int BN = kExBN, STAGES = kExStages;
  constexpr int BK = TF32 ? 32 : 64;        // k extent of one stage (one 128-byte row for K-major operands)
  constexpr int CH = TF32 ? 32 : 64;        // M/N elements per 128-byte chunk of an MN-major operand
  constexpr int UK = TF32 ? 8 : 16;         // k per tcgen05.mma
  constexpr uint32_t CHUNK_BYTES = BK * 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* tiles = smem;
  uint8_t* staging = smem + kExTiles;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kExTiles + kExStg);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (ep.M + kBM - 1) / kBM, num_n = (ep.N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = (num_kb + splits - 1) / splits;
  const int num_items = num_m * num_n * splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int tile = it / splits, ks = it - tile * splits;
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kExStageBytes);
          uint8_t* sa = tiles + size_t(s) * kExStageBytes;
          uint8_t* sb = sa + kExABytes;
          if constexpr (A_MN) {
#pragma unroll
            for (int c = 0; c < kBM / CH; ++c) tma_load_2d(sa + c * CHUNK_BYTES, &tmA, &full[s], m_blk * kBM + c * CH, kb * BK);
          } else {
            tma_load_2d(sa, &tmA, &full[s], kb * BK, m_blk * kBM);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / CH; ++c) tma_load_2d(sb + c * CHUNK_BYTES, &tmB, &full[s], n_blk * BN + c * CH, kb * BK);
          } else {
            tma_load_2d(sb, &tmB, &full[s], kb * BK, n_blk * BN);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_major(uint32_t(FMT), kBM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
        const int tile = it / splits, ks = it - tile * splits;
        const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        mbar_wait(&tempty[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = base + uint32_t(s) * kExStageBytes;
          const uint32_t sb = sa + kExABytes;
          const uint64_t adesc = A_MN ? make_sw128_mnmajor_desc(sa, CHUNK_BYTES) : make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = B_MN ? make_sw128_mnmajor_desc(sb, CHUNK_BYTES) : make_sw128_kmajor_desc(sb);
          constexpr uint64_t a_step = A_MN ? uint64_t((UK * 128) >> 4) : 2ull;   // MN-major: UK k-rows of 128 B; K-major: 32 B
          constexpr uint64_t b_step = B_MN ? uint64_t((UK * 128) >> 4) : 2ull;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_ss<TF32>(d_tmem, adesc + a_step * uint64_t(k), bdesc + b_step * uint64_t(k), idesc, (kb > kb0 || k) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (kb == kb1 - 1) umma_commit(&tfull[as]);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (kb1 <= kb0) umma_commit(&tfull[as]);   // empty split (cannot happen with the host's split choice)
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    const int e = warp - 2;
    const int quad = warp & 3;
    const int half = e >> 2;
    uint8_t* stg = staging + e * 2 * kStageBufBytes;
    uint32_t stg_sel = 0;
    int as = 0; uint32_t aph = 0;
    for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
      const int tile = it / splits;
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      if (!atomic) {
        epilogue_tile<BN>(ep, &tmA, stg, tmem_base + uint32_t(as * BN), &tfull[as], aph, m_blk, n_blk, quad, half, lane, 2, &stg_sel);
      } else {
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        const int row = m_blk * kBM + quad * 32 + lane;
        float* outp = reinterpret_cast<float*>(ep.out);
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 32) {
          const int col_local = half * (BN / 2) + c;
          const int gcol = n_blk * BN + col_local;
          if (gcol >= ep.N) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + uint32_t(as * BN) + (uint32_t(quad * 32) << 16) + uint32_t(col_local), r);
          tmem_ld_wait();
          if (row < ep.M) {
            float* o = outp + static_cast<long long>(row) * ep.ldo + gcol;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (gcol + j < ep.N) atomicAdd(o + j, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// Exact-fp32 reference form (validation mode): same layouts, one output element per thread-register.
__global__ void __launch_bounds__(256)
gemm_ex_simt_kernel(const float* __restrict__ A, long long lda, int a_mn, const float* __restrict__ B, long long ldb, int b_mn,
                    int K, EpiParams ep) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      {
        const int r = a_mn ? (i & 63) : (i >> 4), k = a_mn ? (i >> 6) : (i & 15);
        const int gm = m0 + r, gk = k0 + k;
        As[k][r] = (gm < ep.M && gk < K) ? (a_mn ? A[gk * lda + gm] : A[gm * lda + gk]) : 0.0f;
      }
      {
        const int r = b_mn ? (i & 63) : (i >> 4), k = b_mn ? (i >> 6) : (i & 15);
        const int gn = n0 + r, gk = k0 + k;
        Bs[k][r] = (gn < ep.N && gk < K) ? (b_mn ? B[gk * ldb + gn] : B[gn * ldb + gk]) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= ep.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < ep.N) epi_store_scalar(ep, row, col, acc[i][j]);
    }
  }
}

template <int FMT, bool A_MN, bool B_MN>
static int launch_ex_t(const CUtensorMap& tmA, const CUtensorMap& tmB, int K, int splits, int atomic, const EpiParams& ep,
                       cudaStream_t stream) {
  static DeviceOnce once;
  auto kern = gemm_ex_kernel<FMT, A_MN, B_MN>;
  if (once.first()) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kExSmem)));
  }
  const int items = ((ep.M + kBM - 1) / kBM) * ((ep.N + kExBN - 1) / kExBN) * splits;
  const int ctas = items < num_sms() ? items : num_sms();
  kern<<<ctas, kGemmThreads, kExSmem, stream>>>(tmA, tmB, K, splits, atomic, ep);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

template <int FMT>
static int launch_ex_fmt(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, int K, int splits, int atomic,
                         const EpiParams& ep, cudaStream_t st) {
  if (a_mn && b_mn) return launch_ex_t<FMT, true, true>(tmA, tmB, K, splits, atomic, ep, st);
  if (a_mn) return launch_ex_t<FMT, true, false>(tmA, tmB, K, splits, atomic, ep, st);
  if (b_mn) return launch_ex_t<FMT, false, true>(tmA, tmB, K, splits, atomic, ep, st);
  return launch_ex_t<FMT, false, false>(tmA, tmB, K, splits, atomic, ep, st);
}

// out[M,N] (fp32 / 16-bit) = (accumulate ? out : 0) + A x B with the layouts above.  `accumulate` needs fp32 output.
int launch_gemm_ex(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int in_dtype, int M, int N,
                   int K, void* out, long long ldo, int out_dtype, int accumulate, int impl, int split_k, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  CSVIT_REQUIRE(!accumulate || out_dtype == DT_F32, "gemm_ex: accumulation needs an fp32 output");
  EpiParams ep{};
  ep.out = out; ep.ldo = ldo; ep.out_dtype = out_dtype; ep.act = ACT_NONE;
  ep.M = M; ep.N = N;
  ep.map_mode = ROWMAP_IDENTITY;
  ep.geom = make_geom(1, 1, 1, 0);
  const size_t oes = dtype_size(out_dtype);
  ep.vec_ok = (N % 8 == 0) && ((ldo * oes) % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (K <= 0) {
    if (!accumulate) CSVIT_CUDA(cudaMemset2DAsync(out, ldo * oes, 0, N * oes, M, stream));
    return 0;
  }
  if (impl == GEMM_SIMT) {
    CSVIT_REQUIRE(in_dtype == DT_F32, "SIMT GEMM takes fp32 operands only");
    if (accumulate) { ep.resid = static_cast<const float*>(out); ep.ldr = ldo; }
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    gemm_ex_simt_kernel<<<grid, 256, 0, stream>>>(static_cast<const float*>(A), lda, a_mn, static_cast<const float*>(B), ldb,
                                                  b_mn, K, ep);
    CSVIT_CUDA(cudaGetLastError());
    return 0;
  }
  CSVIT_REQUIRE(in_dtype != DT_F32 || (!a_mn && !b_mn),
                "gemm_ex: kind::tf32 takes K-major operands only (transpose fp32 MN-major operands with csvit_transpose_f32)");
  const int BK = in_dtype == DT_F32 ? 32 : 64;
  const int num_kb = (K + BK - 1) / BK;
  const int tiles = ((M + kBM - 1) / kBM) * ((N + kExBN - 1) / kExBN);
  int splits = split_k;
  if (splits <= 0) {
    splits = 1;
    if (out_dtype == DT_F32 && tiles * 2 <= num_sms() && num_kb >= 16) {
      splits = num_sms() / tiles;            // floor: tiles * splits <= SMs, one full wave (ceil spills a few items into a second wave)
      if (splits > num_kb / 8) splits = num_kb / 8;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > num_kb) splits = num_kb;
  if (splits > 1) {   // no empty splits: shrink so that every split owns at least one k-block
    const int per = (num_kb + splits - 1) / splits;
    splits = (num_kb + per - 1) / per;
  }
  CSVIT_REQUIRE(splits == 1 || out_dtype == DT_F32, "gemm_ex: split-K needs an fp32 output");
  const int atomic = splits > 1 ? 1 : 0;
  if (atomic && !accumulate) CSVIT_CUDA(cudaMemset2DAsync(out, ldo * oes, 0, N * oes, M, stream));
  if (!atomic) {
    if (accumulate) { ep.resid = static_cast<const float*>(out); ep.ldr = ldo; }
    ep.coalesced = (out_dtype == DT_F32 && ep.vec_ok && (N % 32 == 0) && (!ep.resid || ldo % 4 == 0)) ? 1 : 0;
  }
  CUtensorMap tmA, tmB;
  if (a_mn) { if (int e = make_tmap(&tmA, A, lda, K, M, in_dtype, BK, true)) return e; }
  else      { if (int e = make_tmap(&tmA, A, lda, M, K, in_dtype, kBM, true)) return e; }
  if (b_mn) { if (int e = make_tmap(&tmB, B, ldb, K, N, in_dtype, BK, true)) return e; }
  else      { if (int e = make_tmap(&tmB, B, ldb, N, K, in_dtype, kExBN, true)) return e; }
  if (in_dtype == DT_F32) return launch_ex_fmt<2>(a_mn, b_mn, tmA, tmB, K, splits, atomic, ep, stream);
  if (in_dtype == DT_BF16) return launch_ex_fmt<1>(a_mn, b_mn, tmA, tmB, K, splits, atomic, ep, stream);
  return launch_ex_fmt<0>(a_mn, b_mn, tmA, tmB, K, splits, atomic, ep, stream);
}

}  // namespace csvit
