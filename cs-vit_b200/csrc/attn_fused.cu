// Fused shifted-window attention for the narrow Swin stages (C = 128, 256): ONE kernel from the fp32 residual
// stream to the token-ordered attention context,
//     ctx[token] = concat_h softmax( (LN(x) Wq_h^T + bq)(LN(x) Wk_h^T + bk)^T / sqrt(32) + bias_h + shift_mask ) (LN(x) Wv_h^T + bv)
// i.e. layernorm_before + pad + roll(-s) + window_partition + query/key/value + Q K^T + relative-position-bias gather + mask add +
// softmax + P V + head merge + window_reverse + roll(+s) of HF:swin/modeling_swin.py:404-459, 556-582, 604-636, on tcgen05 / TMEM.
// The normalised activations, Q, K, V, the logits and the probabilities never leave the SM: per token the kernel reads 4C bytes
// (fp32 x) and writes 2C (16-bit ctx) where LayerNorm + QKV GEMM + attention kernels moved 6C + 8C + 8C.
//
// Tile = two 49-token windows in one 128-row MMA tile (window A rows 0-48, window B rows 64-112; other rows are zero padding).
//   LN      4 producer warps gather the 98 token rows (closed-form shift/partition map), normalise them in registers and write the
//           16-bit tile straight into the 128-byte-swizzled K-major layout tcgen05 reads; it stays resident for all heads.
//   G(h)    acc[128 x 96] = xn[128 x C] * Wqkv_h[96 x C]^T    (per-head rows q|k|v of the packed weight, streamed by TMA)
//   D(h)    acc -> +bias -> 16 bit -> Q' [128 x 64] (rows of window A carry q in columns 0-31 and zeros in 32-63, window B the
//           other way round), K' [64 keys x 64] = [K_A | K_B], V' = [V_A | V_B] (the same bytes read as an MN-major operand)
//   S(h)    S'[128 x 64] = Q' K'^T + I Bias_h^T : the block structure of Q' makes row r meet only the keys of its own window, so
//           two windows share one M = 128 MMA with N = 64; the relative-position bias (pre-multiplied by log2 e, -30000 in the
//           padding key columns) enters as a second K block against a constant one-hot operand - no bias loads in the softmax.
//   X(h)    one thread per row: 56 logits from TMEM, shift mask from a closed-form bit mask (only windows on the last window
//           row / column), exp2, unnormalised P in 16 bit written over Q' (dead once S(h) has completed)
//   PV(h)   O'[128 x 64] = P' V' : row r finds its window's output in columns 32 (r / 64) .. +31
//   E(h)    O' / rowsum -> 16 bit -> ctx rows in TOKEN order (window_reverse + un-shift folded into the store address)
// Roles: warp 0 TMA producer (+ L2 prefetch of the next tiles' rows), warp 1 MMA issuer, warps 2-5 / 6-9 two softmax warpgroups
// that take alternate heads (so one group's exponentials overlap the other's TMEM drain and the MMAs of both), warps 10-13
// LayerNorm producers.  All per-head resources (accumulator, Q'K'V' buffers, S', O') are double-buffered by head parity; the MMA
// issue order G(s), S(s-1), PV(s-2) runs continuously across tile boundaries.  TMEM: 2 x 96 + 2 x 64 + 2 x 64 = 448 columns.
#include <type_traits>

#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"

namespace csvit {

constexpr int FA_THREADS = 448;
constexpr int FA_L = 49;
constexpr int FA_ROWS = 2 * FA_L;                   // real token rows of a tile
constexpr uint32_t FA_SLAB = 128 * 128;             // 128 rows x 64 16-bit columns
constexpr uint32_t FA_STAGE = 96 * 128;             // ring stage: one head's 96 weight rows x 64 columns (or a 64 x 64 bias operand)
constexpr int FA_NST = 4;
constexpr uint32_t FA_QP = 0, FA_K = 16384, FA_V = 24576, FA_HB = 32768;   // per-parity head buffer: Q'/P', K', V'
constexpr uint32_t FA_TM_ACC = 0, FA_TM_S = 192, FA_TM_O = 320;
constexpr float FA_MASK_LOG2 = -100.0f * 1.4426950408889634f;

template <int C>
struct FaCfg {
  static constexpr int KB = C / 64;
  static constexpr int HEADS = C / 32;
  static constexpr int XNB = C <= 128 ? 2 : 1;      // resident LayerNorm tiles (double-buffered where shared memory allows)
  static constexpr uint32_t XN_TILE = KB * FA_SLAB;
  static constexpr uint32_t XN_OFF = 0;
  static constexpr uint32_t RING_OFF = XNB * XN_TILE;
  static constexpr uint32_t HB_OFF = RING_OFF + FA_NST * FA_STAGE;
  static constexpr uint32_t ID_OFF = HB_OFF + 2 * FA_HB;
  static constexpr uint32_t BAR_OFF = ID_OFF + FA_SLAB;
  static constexpr size_t SMEM = 1024 + size_t(BAR_OFF) + 512;
};

struct FaParams {
  const float* x;        // fp32 residual stream [B*N, C]
  const float* gamma;
  const float* beta;
  float eps;
  const float* bqkv;     // [heads * 96] fp32, per head q | k | v, q part pre-multiplied by qscale
  void* ctx;             // 16-bit [B*N, C], token order
  int num_windows;       // B * nW
  int nW;
  float qscale;          // log2(e) / sqrt(32)
  WinGeom g;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ uint64_t fa_mnmajor_desc(uint32_t smem_addr) {
  // MN-major operand of one 128-byte MN chunk (64 elements): rows = k, 8-row swizzle atoms of 1024 B (layout of gemm_ex.cu)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;    // LBO: next MN chunk (unused, N = 64 is one chunk)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: next group of 8 k-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// FMT: 0 = fp16, 1 = bf16.
template <int FMT, int C>
__global__ void __launch_bounds__(FA_THREADS, 1)
swin_attn_fused_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmB, FaParams p) {
  using Cfg = FaCfg<C>;
  using T16 = typename std::conditional<FMT == 1, __nv_bfloat16, __half>::type;
  constexpr int KB = Cfg::KB, HEADS = Cfg::HEADS, XNB = Cfg::XNB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* w_full = bars;               // [FA_NST]
  uint64_t* w_empty = bars + FA_NST;     // [FA_NST]
  uint64_t* xn_full = bars + 2 * FA_NST; // [2]
  uint64_t* xn_empty = xn_full + 2;      // [2]
  uint64_t* acc_full = xn_full + 4;      // [2] by head parity, like everything below
  uint64_t* acc_empty = xn_full + 6;
  uint64_t* qkv_full = xn_full + 8;
  uint64_t* s_full = xn_full + 10;
  uint64_t* p_full = xn_full + 12;
  uint64_t* o_full = xn_full + 14;
  uint64_t* o_empty = xn_full + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xn_full + 18);
  volatile int* ln_done = reinterpret_cast<volatile int*>(tmem_slot + 1);   // tiles whose LayerNorm is complete (prefetch pacing)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.num_windows + 1) >> 1;
  const int my_tiles = (num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  // zero everything the MMAs read but no producer rewrites (padding rows, the off-window halves), then the one-hot operand
  for (uint32_t i = threadIdx.x; i < Cfg::RING_OFF / 16; i += FA_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = threadIdx.x; i < (2 * FA_HB + FA_SLAB) / 16; i += FA_THREADS)
    reinterpret_cast<uint4*>(smem + Cfg::HB_OFF)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmB);
    for (int s = 0; s < FA_NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&xn_full[b], 4); mbar_init(&xn_empty[b], 1);
      mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4);
      mbar_init(&qkv_full[b], 4); mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 4); mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 4);
    }
    *ln_done = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  __syncthreads();
  if (threadIdx.x < 128) {   // one-hot rows: I[r][r % 64] = 1 (fp16) for the 49 real slots of each window
    const int r = threadIdx.x, i = r & 63;
    if (i < FA_L)
      *reinterpret_cast<uint16_t*>(smem + Cfg::ID_OFF + r * 128 + (((i >> 3) ^ (r & 7)) << 4) + (i & 7) * 2) = 0x3C00u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: per head the KB weight slabs, then (one step later, as the MMA order needs it) the bias operand
      int s = 0; uint32_t ph = 0;
      auto push = [&](const CUtensorMap* tm, int col, int row, uint32_t bytes) {
        mbar_wait(&w_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&w_full[s], bytes);
        tma_load_2d(smem + Cfg::RING_OFF + size_t(s) * FA_STAGE, tm, &w_full[s], col, row);
        if (++s == FA_NST) { s = 0; ph ^= 1u; }
      };
      const int total = my_tiles * HEADS;
      for (int gs = 0; gs <= total; ++gs) {
        if (gs < total) {
          const int h = gs % HEADS;
          for (int kb = 0; kb < KB; ++kb) push(&tmW, kb * 64, h * 96, 96 * 128);
        }
        if (gs >= 1) push(&tmB, 0, ((gs - 1) % HEADS) * 64, 64 * 128);
      }
    } else {
      // ---------------- idle lanes: pull the rows of the tiles ahead into L2, paced by the LayerNorm warps' progress
      auto prefetch_tile = [&](int ti) {
        if (ti >= my_tiles) return;
        const int t = int(blockIdx.x) + ti * int(gridDim.x);
        for (int q = lane - 1; q < FA_ROWS; q += 31) {
          const int wg = 2 * t + q / FA_L, i = q % FA_L;
          if (wg >= p.num_windows) continue;
          const int b = wg / p.nW, w = wg - b * p.nW;
          const char* src = reinterpret_cast<const char*>(p.x + (static_cast<long long>(b) * p.g.N + win_row_to_token(p.g, w * FA_L + i)) * C);
#pragma unroll
          for (int o = 0; o < C * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
        }
      };
      for (int d = 1; d <= XNB; ++d) prefetch_tile(d);
      for (int ti = 1; ti + XNB < my_tiles; ++ti) {
        while (*ln_done < ti) __nanosleep(256);
        prefetch_tile(ti + XNB);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_g = make_idesc(uint32_t(FMT), 128, 96);
      constexpr uint32_t idesc_s = make_idesc(uint32_t(FMT), 128, 64);
      constexpr uint32_t idesc_b = make_idesc(0u, 128, 64);                                        // one-hot x bias: always fp16
      constexpr uint32_t idesc_o = make_idesc(uint32_t(FMT), 128, 64) | (1u << 16);                // V' is MN-major
      int s = 0; uint32_t ph = 0;
      uint32_t nG[2] = {0, 0}, nS[2] = {0, 0}, nP[2] = {0, 0};
      const int total = my_tiles * HEADS;
      uint32_t xn_addr = 0;
      for (int gs = 0; gs < total + 2; ++gs) {
        if (gs < total) {
          const int h = gs % HEADS, ti = gs / HEADS, b = h & 1;
          const int xb = XNB == 2 ? (ti & 1) : 0;
          if (h == 0) {
            mbar_wait(&xn_full[xb], XNB == 2 ? ((ti >> 1) & 1) : (ti & 1));
            xn_addr = base + Cfg::XN_OFF + uint32_t(xb) * Cfg::XN_TILE;
          }
          mbar_wait(&acc_empty[b], (nG[b] & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + FA_TM_ACC + uint32_t(b * 96);
#pragma unroll 1
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(xn_addr + uint32_t(kb) * FA_SLAB);
            const uint64_t bdesc = make_sw128_kmajor_desc(base + Cfg::RING_OFF + uint32_t(s) * FA_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_g, (kb | k) ? 1u : 0u);
            umma_commit(&w_empty[s]);
            if (++s == FA_NST) { s = 0; ph ^= 1u; }
          }
          umma_commit(&acc_full[b]);
          ++nG[b];
          if (h == HEADS - 1) umma_commit(&xn_empty[xb]);     // the tile's last projection issued: xn may be overwritten once it completes
        }
        if (gs >= 1 && gs <= total) {
          const int b = (gs - 1) & 1;     // HEADS is even: head parity = global-step parity
          const uint32_t hb = base + Cfg::HB_OFF + uint32_t(b) * FA_HB;
          mbar_wait(&qkv_full[b], nS[b] & 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + FA_TM_S + uint32_t(b * 64);
          const uint64_t qdesc = make_sw128_kmajor_desc(hb + FA_QP), kdesc = make_sw128_kmajor_desc(hb + FA_K);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
          mbar_wait(&w_full[s], ph);
          tc_fence_after();
          const uint64_t idesc_a = make_sw128_kmajor_desc(base + Cfg::ID_OFF);
          const uint64_t bdesc = make_sw128_kmajor_desc(base + Cfg::RING_OFF + uint32_t(s) * FA_STAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, idesc_a + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_b, 1u);
          umma_commit(&w_empty[s]);
          if (++s == FA_NST) { s = 0; ph ^= 1u; }
          umma_commit(&s_full[b]);
          ++nS[b];
        }
        if (gs >= 2) {
          const int b = gs & 1;
          const uint32_t hb = base + Cfg::HB_OFF + uint32_t(b) * FA_HB;
          mbar_wait(&p_full[b], nP[b] & 1u);
          mbar_wait(&o_empty[b], (nP[b] & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + FA_TM_O + uint32_t(b * 64);
          const uint64_t pdesc = make_sw128_kmajor_desc(hb + FA_QP);
          const uint64_t vdesc = fa_mnmajor_desc(hb + FA_V);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, pdesc + uint64_t(2 * k), vdesc + uint64_t(128 * k), idesc_o, k ? 1u : 0u);
          umma_commit(&o_full[b]);
          ++nP[b];
        }
      }
    }
  } else if (warp < 10) {
    // ---------------- softmax warpgroups: group g takes the heads of parity g ----------------
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int wdx = r >> 6, j = r & 63;      // window of the pair, slot inside it
    const bool bf = FMT == 1;
    uint8_t* hb = smem + Cfg::HB_OFF + uint32_t(g) * FA_HB;
    const uint32_t lane_bits = uint32_t(quad * 32) << 16;
    const int nWy = p.g.H / p.g.ws;
    T16* ctx = static_cast<T16*>(p.ctx);
    const int per_group = my_tiles * (HEADS / 2);
    long long tok_off = -1;
    uint32_t dm_lo = 0, dm_hi = 0;           // shift mask of this row: bit c set = key slot c lies in another region
    for (int n = 0; n < per_group; ++n) {
      const int gh = 2 * n + g;
      const int ti = gh / HEADS, h = gh - ti * HEADS;
      const uint32_t ph = uint32_t(n) & 1u;
      if (h == g) {   // first head of a tile for this group: where the row goes, and its mask
        const int t = int(blockIdx.x) + ti * int(gridDim.x);
        const int wg = 2 * t + wdx;
        tok_off = -1; dm_lo = dm_hi = 0;
        if (j < FA_L && wg < p.num_windows) {
          const int b = wg / p.nW, w = wg - b * p.nW;
          tok_off = (static_cast<long long>(b) * p.g.N + win_row_to_token(p.g, w * FA_L + j)) * C;
          if (p.g.shift > 0) {
            const int wy = w / p.g.nWx, wx = w - wy * p.g.nWx;
            const int iy = j / 7, ix = j - iy * 7, th = p.g.ws - p.g.shift;
            unsigned long long dm = 0;
            if (wy == nWy - 1) {      // key slots whose row side (iy < th) differs from mine
              const unsigned long long rows_lo = (1ull << (7 * th)) - 1ull;
              dm |= (iy < th) ? ~rows_lo : rows_lo;
            }
            if (wx == p.g.nWx - 1) {
              unsigned long long cols_lo = 0;
              for (int c = 0; c < FA_L; ++c) if (c % 7 < th) cols_lo |= 1ull << c;
              dm |= (ix < th) ? ~cols_lo : cols_lo;
            }
            dm &= (1ull << FA_L) - 1ull;
            dm_lo = uint32_t(dm); dm_hi = uint32_t(dm >> 32);
          }
        }
      }
      // ---- D(h): projection accumulator -> Q' / K' / V'
      {
        mbar_wait(&acc_full[g], ph);
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_bits + FA_TM_ACC + uint32_t(g * 96);
        const float4* bp = reinterpret_cast<const float4*>(p.bqkv + h * 96);
        uint32_t rq[32], rk[32];
        tmem_ld_32x32(ta, rq);
        tmem_ld_32x32(ta + 32u, rk);
        tmem_ld_wait();
        uint32_t pq[16], pk[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b0 = __ldg(bp + c), b1 = __ldg(bp + 8 + c);
          pq[2 * c] = pack16(bf, fmaf(__uint_as_float(rq[4 * c]), p.qscale, b0.x), fmaf(__uint_as_float(rq[4 * c + 1]), p.qscale, b0.y));
          pq[2 * c + 1] = pack16(bf, fmaf(__uint_as_float(rq[4 * c + 2]), p.qscale, b0.z), fmaf(__uint_as_float(rq[4 * c + 3]), p.qscale, b0.w));
          pk[2 * c] = pack16(bf, __uint_as_float(rk[4 * c]) + b1.x, __uint_as_float(rk[4 * c + 1]) + b1.y);
          pk[2 * c + 1] = pack16(bf, __uint_as_float(rk[4 * c + 2]) + b1.z, __uint_as_float(rk[4 * c + 3]) + b1.w);
        }
        uint32_t rv[32];
        tmem_ld_32x32(ta + 64u, rv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[g]);   // accumulator drained: the projection of head h + 2 may overwrite it
        uint32_t pv[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b2 = __ldg(bp + 16 + c);
          pv[2 * c] = pack16(bf, __uint_as_float(rv[4 * c]) + b2.x, __uint_as_float(rv[4 * c + 1]) + b2.y);
          pv[2 * c + 1] = pack16(bf, __uint_as_float(rv[4 * c + 2]) + b2.z, __uint_as_float(rv[4 * c + 3]) + b2.w);
        }
        // (the buffers of this parity are free: E(h - 2) of this warp has seen PV(h - 2) complete)
        uint8_t* qrow = hb + FA_QP + r * 128;
        uint8_t* krow = hb + FA_K + j * 128;
        uint8_t* vrow = hb + FA_V + j * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t own = uint32_t(((wdx * 4 + c) ^ (r & 7)) << 4), other = uint32_t((((1 - wdx) * 4 + c) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(qrow + own) = make_uint4(pq[4 * c], pq[4 * c + 1], pq[4 * c + 2], pq[4 * c + 3]);
          *reinterpret_cast<uint4*>(qrow + other) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(krow + own) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(vrow + own) = make_uint4(pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qkv_full[g]);
      }
      // ---- X(h): logits -> unnormalised probabilities (log2 domain), written over Q'
      float rsum;
      {
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        const uint32_t ts = tmem_base + lane_bits + FA_TM_S + uint32_t(g * 64);
        uint32_t s0[32], s1[16], s2[8];
        tmem_ld_32x32(ts, s0);
        tmem_ld_32x16(ts + 32u, s1);
        tmem_ld_32x8(ts + 48u, s2);
        tmem_ld_wait();
        float sv[56];
#pragma unroll
        for (int c = 0; c < 32; ++c) sv[c] = __uint_as_float(s0[c]);
#pragma unroll
        for (int c = 0; c < 16; ++c) sv[32 + c] = __uint_as_float(s1[c]);
#pragma unroll
        for (int c = 0; c < 8; ++c) sv[48 + c] = __uint_as_float(s2[c]);
        if ((dm_lo | dm_hi) != 0u) {
#pragma unroll
          for (int c = 0; c < FA_L; ++c) {
            const uint32_t bit = c < 32 ? (dm_lo >> c) & 1u : (dm_hi >> (c - 32)) & 1u;
            sv[c] += bit ? FA_MASK_LOG2 : 0.0f;
          }
        }
        float mx0 = sv[0], mx1 = sv[1];
#pragma unroll
        for (int c = 2; c < 56; c += 2) { mx0 = fmaxf(mx0, sv[c]); mx1 = fmaxf(mx1, sv[c + 1]); }
        const float mx = fmaxf(mx0, mx1);
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pp[28];
#pragma unroll
        for (int c = 0; c < 28; ++c) {
          const float e0 = fa_exp2(sv[2 * c] - mx), e1 = fa_exp2(sv[2 * c + 1] - mx);
          sum0 += e0; sum1 += e1;
          pp[c] = pack16(bf, e0, e1);
        }
        rsum = sum0 + sum1;
        uint8_t* prow = hb + FA_QP + r * 128;
#pragma unroll
        for (int c = 0; c < 7; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        *reinterpret_cast<uint4*>(prow + ((7 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---- E(h): O' / rowsum -> token-ordered context
      {
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + lane_bits + FA_TM_O + uint32_t(g * 64 + wdx * 32), o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[g]);
        if (tok_off >= 0) {
          const float inv = 1.0f / rsum;
          uint4* dst = reinterpret_cast<uint4*>(ctx + tok_off + h * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack16(bf, __uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                                pack16(bf, __uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                                pack16(bf, __uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                                pack16(bf, __uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
        }
      }
    }
  } else {
    // ---------------- LayerNorm producers: gathered fp32 rows -> normalised 16-bit tile in the MMA layout ----------------
    constexpr int LPR = C >= 256 ? 32 : C / 8;        // lanes per row (each lane owns 8-column chunks)
    constexpr int RPP = 32 / LPR;                     // rows per warp pass
    constexpr int CPL = C / 8 / LPR;                  // chunks per lane
    constexpr int NPASS = (FA_ROWS + 4 * RPP - 1) / (4 * RPP);
    constexpr int GP = 7;                             // passes whose loads are in flight together
    const int ln = warp - 10;
    const int sub = lane / LPR, lr = lane % LPR;
    float gam[CPL][8], bet[CPL][8];
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = lr + LPR * t;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c * 8 + 4));
      gam[t][0] = g0.x; gam[t][1] = g0.y; gam[t][2] = g0.z; gam[t][3] = g0.w; gam[t][4] = g1.x; gam[t][5] = g1.y; gam[t][6] = g1.z; gam[t][7] = g1.w;
      bet[t][0] = b0.x; bet[t][1] = b0.y; bet[t][2] = b0.z; bet[t][3] = b0.w; bet[t][4] = b1.x; bet[t][5] = b1.y; bet[t][6] = b1.z; bet[t][7] = b1.w;
    }
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int t = int(blockIdx.x) + ti * int(gridDim.x);
      const int xb = XNB == 2 ? (ti & 1) : 0;
      uint8_t* xn = smem + Cfg::XN_OFF + uint32_t(xb) * Cfg::XN_TILE;
      bool waited = false;
#pragma unroll 1
      for (int p0 = 0; p0 < NPASS; p0 += GP) {
        float v[GP][CPL][8];
        int rowq[GP];
#pragma unroll
        for (int pp = 0; pp < GP; ++pp) {
          const int q = ((p0 + pp) * 4 + ln) * RPP + sub;
          const int wg = 2 * t + q / FA_L, i = q % FA_L;
          const bool ok = (p0 + pp) < NPASS && q < FA_ROWS && wg < p.num_windows;
          rowq[pp] = ok ? (q / FA_L) * 64 + i : -1;
          if (ok) {
            const int b = wg / p.nW, w = wg - b * p.nW;
            const float* xr = p.x + (static_cast<long long>(b) * p.g.N + win_row_to_token(p.g, w * FA_L + i)) * C;
#pragma unroll
            for (int tt = 0; tt < CPL; ++tt) {
              const int c = lr + LPR * tt;
              const float4 a0 = *reinterpret_cast<const float4*>(xr + c * 8), a1 = *reinterpret_cast<const float4*>(xr + c * 8 + 4);
              v[pp][tt][0] = a0.x; v[pp][tt][1] = a0.y; v[pp][tt][2] = a0.z; v[pp][tt][3] = a0.w;
              v[pp][tt][4] = a1.x; v[pp][tt][5] = a1.y; v[pp][tt][6] = a1.z; v[pp][tt][7] = a1.w;
            }
          } else {
#pragma unroll
            for (int tt = 0; tt < CPL; ++tt)
#pragma unroll
              for (int e = 0; e < 8; ++e) v[pp][tt][e] = 0.f;
          }
        }
        if (!waited) {   // the raw rows do not depend on the tile buffer being free: they were requested first
          mbar_wait(&xn_empty[xb], (XNB == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1u);
          waited = true;
        }
#pragma unroll
        for (int pp = 0; pp < GP; ++pp) {
          float sum = 0.f;
#pragma unroll
          for (int tt = 0; tt < CPL; ++tt)
            sum += ((v[pp][tt][0] + v[pp][tt][1]) + (v[pp][tt][2] + v[pp][tt][3])) + ((v[pp][tt][4] + v[pp][tt][5]) + (v[pp][tt][6] + v[pp][tt][7]));
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float mean = sum * (1.0f / float(C));
          float sq = 0.f;
#pragma unroll
          for (int tt = 0; tt < CPL; ++tt)
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = v[pp][tt][e] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          const float rstd = rsqrtf(sq * (1.0f / float(C)) + p.eps);
          const int rr = rowq[pp];
          if (rr >= 0) {
#pragma unroll
            for (int tt = 0; tt < CPL; ++tt) {
              const int c = lr + LPR * tt;
              const float* w = v[pp][tt];
              uint4 pk;
              pk.x = Half16<T16>::pack((w[0] - mean) * rstd * gam[tt][0] + bet[tt][0], (w[1] - mean) * rstd * gam[tt][1] + bet[tt][1]);
              pk.y = Half16<T16>::pack((w[2] - mean) * rstd * gam[tt][2] + bet[tt][2], (w[3] - mean) * rstd * gam[tt][3] + bet[tt][3]);
              pk.z = Half16<T16>::pack((w[4] - mean) * rstd * gam[tt][4] + bet[tt][4], (w[5] - mean) * rstd * gam[tt][5] + bet[tt][5]);
              pk.w = Half16<T16>::pack((w[6] - mean) * rstd * gam[tt][6] + bet[tt][6], (w[7] - mean) * rstd * gam[tt][7] + bet[tt][7]);
              *reinterpret_cast<uint4*>(xn + (c >> 3) * FA_SLAB + rr * 128 + (((c & 7) ^ (rr & 7)) << 4)) = pk;
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&xn_full[xb]);
        if (ln == 0) *ln_done = ti + 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int FMT, int C>
static int launch_fa(const CUtensorMap& tmW, const CUtensorMap& tmB, const FaParams& p, cudaStream_t stream) {
  using Cfg = FaCfg<C>;
  auto kern = swin_attn_fused_kernel<FMT, C>;
  CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
  const int tiles = (p.num_windows + 1) / 2;
  const int ctas = tiles < num_sms() ? tiles : num_sms();
  kern<<<ctas, FA_THREADS, Cfg::SMEM, stream>>>(tmW, tmB, p);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_swin_attn_fused(const float* x, const float* gamma, const float* beta, float eps, const void* wqkv_h, const float* bqkv_h,
                           const void* bias_op, void* ctx, int dtype, int B, int H, int W, int C, int heads, int ws, int shift,
                           cudaStream_t stream) {
  CSVIT_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "swin_attn_fused: 16-bit operand formats only");
  CSVIT_REQUIRE(C == 128 || C == 256, "swin_attn_fused: C=%d not in {128, 256}", C);
  CSVIT_REQUIRE(ws == 7 && C == heads * 32, "swin_attn_fused: window 7 / head_dim 32 only (ws=%d C=%d heads=%d)", ws, C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "swin_attn_fused: bad geometry %dx%d shift %d", H, W, shift);
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bqkv_h) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
                "swin_attn_fused: operands must be 16-byte aligned");
  const int nW = (H / ws) * (W / ws);
  const long long windows = static_cast<long long>(B) * nW;
  if (windows <= 0) return 0;
  CSVIT_REQUIRE(windows < (1ll << 30), "swin_attn_fused: too many windows");
  FaParams p{};
  p.x = x; p.gamma = gamma; p.beta = beta; p.eps = eps; p.bqkv = bqkv_h; p.ctx = ctx;
  p.num_windows = static_cast<int>(windows); p.nW = nW;
  p.qscale = 1.4426950408889634f * 0.17677669529663687f;
  p.g = make_geom(H, W, ws, shift);
  CUtensorMap tmW, tmB;
  if (int e = make_tmap(&tmW, wqkv_h, C, 3ll * C, C, dtype, 96, true)) return e;
  if (int e = make_tmap(&tmB, bias_op, 64, 64ll * heads, 64, DT_F16, 64, true)) return e;
  const bool bf = dtype == DT_BF16;
  if (C == 128) return bf ? launch_fa<1, 128>(tmW, tmB, p, stream) : launch_fa<0, 128>(tmW, tmB, p, stream);
  return bf ? launch_fa<1, 256>(tmW, tmB, p, stream) : launch_fa<0, 256>(tmW, tmB, p, stream);
}

}  // namespace csvit
