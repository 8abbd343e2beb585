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
//           16-bit tile straight into the 128-byte-swizzled K-major layout tcgen05 reads; it stays resident for all heads.  Only
//           (x - mean) * rstd is formed here: LayerNorm's gamma scales the weight columns and beta moves into the biases at
//           packing time (W' = W diag(gamma), b' = b + W beta, exact in fp32 before the 16-bit rounding of W').
//   G(h)    acc[128 x 96] = xn[128 x C] * Wqkv_h[96 x C]^T    (per-head rows q|k|v of the packed weight, streamed by TMA)
//   D(h)    acc -> 16 bit -> Q' [128 x 64] (rows of window A carry q in columns 0-31 and zeros in 32-63, window B the other way
//           round), K' [64 keys x 64] = [K_A | K_B], V' = [V_A | V_B] (the same bytes read as an MN-major operand).  Only the
//           query bias is added here: the key bias shifts every logit of a row by the same q.bk and cancels in the softmax
//           (exactly), the value bias commutes with the row-stochastic P and is added to the normalised output in E.
//   S(h)    S'[128 x 64] = Q' K'^T : the block structure of Q' makes row r meet only the keys of its own window, so two windows
//           share one M = 128 MMA with N = 64.  (Every M = 128, K = 16 MMA step with A in shared memory costs >= 74 cycles whatever
//           N <= 128 is - tools/cuda/mma_rate.cu - so the tensor pipe is paid per k-step: a bias-by-one-hot-MMA form was dropped.)
//   X(h)    one thread per row: 49 logits from TMEM + the row's relative-position bias (fp16 table row streamed per head by a bulk
//           copy, pre-multiplied by log2 e, added with mixed-precision FHADD) + shift mask from a closed-form bit mask (only windows
//           on the last window row / column), exp2, unnormalised P in 16 bit written over Q' (dead once S(h) has completed)
//   PV(h)   O'[128 x 64] = P' V' (written over S' in TMEM): row r finds its window's output in columns 32 (r / 64) .. +31
//   E(h)    O' / rowsum + bv -> 16 bit -> ctx rows in TOKEN order (window_reverse + un-shift folded into the store address)
// Roles: warp 0 TMA producer, warp 1 projection (G) issuer, warp 2 attention (S, PV) issuer (convergent warps, one elected lane
// issues), warp 3 L2 prefetch of the next tiles' rows, warps 4-15 three softmax warpgroups, warps 16-23 LayerNorm producers (setmaxnreg moves registers to the softmax groups).  A head's chain G -> D -> S -> X -> PV -> E crosses the tensor pipe three times, so THREE
// heads are in flight: global head gh uses slot gh % 3 (accumulator, Q'K'V' buffers, S'/O' columns) and warpgroup gh % 3,
// continuously across tile boundaries.  Two issuer warps: the projections run ahead on their own, so a PV or S whose operands
// are ready is never queued behind a projection's weight waits.
// TMEM: 3 slots x (96 accumulator + 64 S'/O') = 480 columns.
#include <type_traits>

#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"
#include "attn_common.cuh"

namespace csvit {

constexpr int FA_THREADS = 768;                    // 6 warpgroups: {TMA, MMA, 2 x L2 prefetch}, 3 x softmax, 2 x LayerNorm
constexpr int FA_SM_WARP0 = 4, FA_LN_WARP0 = 16, FA_LN_WARPS = 8;
constexpr int FA_NSLOT = 3;
constexpr int FA_ROWS = 2 * FA_L;                   // real token rows of a tile
constexpr uint32_t FA_SLAB = 128 * 128;             // 128 rows x 64 16-bit columns
constexpr uint32_t FA_STAGE = 96 * 128;             // ring stage: one head's 96 weight rows x 64 columns
constexpr int FA_NBS = 2;
constexpr int FA_NST = 4;
constexpr uint32_t FA_QP = 0, FA_K = 16384, FA_V = 24576, FA_HB = 32768;   // per-parity head buffer: Q'/P', K', V'
constexpr uint32_t FA_TM_SLOT = 160, FA_TM_S = 96;            // TMEM columns per slot; S' / O' offset inside it
constexpr uint32_t FA_TAB = FA_K + 49 * 128;                  // the 15 padding key rows of K' (never read by a consumer): bias tables

template <int C>
struct FaCfg {
  static constexpr int KB = C / 64;
  static constexpr int HEADS = C / 32;
  static constexpr int XNB = C <= 128 ? 2 : 1;      // resident LayerNorm tiles (double-buffered where shared memory allows)
  static constexpr uint32_t XN_TILE = KB * FA_SLAB;
  static constexpr uint32_t XN_OFF = 0;
  static constexpr uint32_t RING_OFF = XNB * XN_TILE;
  static constexpr uint32_t HB_OFF = RING_OFF + FA_NST * FA_STAGE;
  static constexpr uint32_t BIAS_OFF = HB_OFF + FA_NSLOT * FA_HB;
  static constexpr uint32_t BAR_OFF = BIAS_OFF + FA_NBS * FA_BIAS_STAGE;
  static constexpr size_t SMEM = 1024 + size_t(BAR_OFF) + 512;
};

struct FaParams {
  const float* x;        // fp32 residual stream [B*N, C]
  float eps;
  const float* bqkv;     // [heads * 96] fp32, per head q | k | v, q part pre-multiplied by qscale (the k part is not used)
  const void* bias;      // fp16 [heads][49][56]: log2(e) * relative position bias of (query slot, key slot)
  void* ctx;             // 16-bit [B*N, C], token order
  int num_windows;       // B * nW
  int nW;
  float qscale;          // log2(e) / sqrt(32)
  WinGeom g;
};

#ifdef CSVIT_FA_TRACE
// Development aid (not built by default): clock64 timestamps of block 0's pipeline events, read back with csvit_debug_fa_trace.
__device__ long long fa_trace_buf[16][256];
#define FA_TR(kind, idx) do { if (blockIdx.x == 0 && (idx) < 256) fa_trace_buf[kind][idx] = clock64(); } while (0)
#else
#define FA_TR(kind, idx) do { } while (0)
#endif

// FMT: 0 = fp16, 1 = bf16.
template <int FMT, int C>
__global__ void __launch_bounds__(FA_THREADS, 1)
swin_attn_fused_kernel(const __grid_constant__ CUtensorMap tmW, FaParams p) {
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
  uint64_t* acc_full = xn_full + 4;      // [FA_NSLOT] by head slot, like everything below
  uint64_t* acc_empty = acc_full + FA_NSLOT;
  uint64_t* qkv_full = acc_empty + FA_NSLOT;
  uint64_t* s_full = qkv_full + FA_NSLOT;
  uint64_t* p_full = s_full + FA_NSLOT;
  uint64_t* o_full = p_full + FA_NSLOT;
  uint64_t* b_full = o_full + FA_NSLOT;  // [FA_NBS] bias ring
  uint64_t* b_empty = b_full + FA_NBS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + FA_NBS);
  volatile int* ln_done = reinterpret_cast<volatile int*>(tmem_slot + 1);   // tiles whose LayerNorm is complete (prefetch pacing)
  float* bq_tab = reinterpret_cast<float*>(smem + Cfg::HB_OFF + FA_TAB);            // [C] query bias (pre-scaled), slot 0's K' padding
  float* bv_tab = reinterpret_cast<float*>(smem + Cfg::HB_OFF + FA_HB + FA_TAB);    // [C] value bias, slot 1's K' padding
  static_assert(C * 4 <= 15 * 128, "bias table does not fit the padding key rows");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.num_windows + 1) >> 1;
  const int my_tiles = (num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int total = my_tiles * HEADS;

  // zero everything the MMAs read but no producer rewrites (padding rows, the off-window halves), then the one-hot operand
  for (uint32_t i = threadIdx.x; i < Cfg::RING_OFF / 16; i += FA_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = threadIdx.x; i < (FA_NSLOT * FA_HB) / 16; i += FA_THREADS)
    reinterpret_cast<uint4*>(smem + Cfg::HB_OFF)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmW);
    for (int s = 0; s < FA_NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&xn_full[b], FA_LN_WARPS); mbar_init(&xn_empty[b], 1); }
    for (int b = 0; b < FA_NBS; ++b) { mbar_init(&b_full[b], 1); mbar_init(&b_empty[b], 4); }
    for (int b = 0; b < FA_NSLOT; ++b) {
      mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4);
      mbar_init(&qkv_full[b], 4); mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 4); mbar_init(&o_full[b], 1);
    }
    *ln_done = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += FA_THREADS) {
    bq_tab[c] = p.bqkv[(c >> 5) * 96 + (c & 31)];
    bv_tab[c] = p.bqkv[(c >> 5) * 96 + 64 + (c & 31)];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();      // PDL: the next kernel's prologue may overlap this kernel's tail ...
  griddep_wait();        // ... and this kernel touches global memory only after its predecessors have completed

  if (warp < FA_SM_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      int s = 0; uint32_t ph = 0;
      // weights in head order on the ring; one head's bias rows per bulk copy on the bias ring, as far ahead as the rings allow
      int wq = 0, bq = 0;
      while (wq < total * KB || bq < total) {
        bool moved = false;
        if (wq < total * KB && mbar_test_wait(&w_empty[s], ph ^ 1u)) {
          const int h = (wq / KB) % HEADS, kb = wq % KB;
          mbar_arrive_expect_tx(&w_full[s], 96 * 128);
          tma_load_2d(smem + Cfg::RING_OFF + size_t(s) * FA_STAGE, &tmW, &w_full[s], kb * 64, h * 96);
          if (wq < 128) FA_TR(15, 128 + wq);
          if (++s == FA_NST) { s = 0; ph ^= 1u; }
          ++wq; moved = true;
        }
        if (bq < total) {
          const int bs = bq % FA_NBS;
          if (mbar_test_wait(&b_empty[bs], ((uint32_t(bq) / FA_NBS) & 1u) ^ 1u)) {
            mbar_arrive_expect_tx(&b_full[bs], FA_BIAS_BYTES);
            fa_bulk_load(smem + Cfg::BIAS_OFF + uint32_t(bs) * FA_BIAS_STAGE, static_cast<const char*>(p.bias) + size_t(bq % HEADS) * FA_BIAS_BYTES,
                         FA_BIAS_BYTES, &b_full[bs]);
            ++bq; moved = true;
          }
        }
        if (!moved) __nanosleep(256);      // (both rings run two heads ahead: a slower poll costs nothing and frees issue slots)
      }
    }
  } else if (warp == 1) {
    // ---------------- projection issuer: G(gh) for every head in order, as far ahead as the accumulator slots and the weight ring
    // allow (the whole warp runs the loop and waits, one elected lane issues).  S and PV have an issuer warp of their own: with one
    // in-order issuer for all three (round 2, first version) a ready PV waited ~2900 cycles behind the next head's S and a
    // projection whose issue takes 1000-1700 cycles when the warp shares its scheduler with five busy warps (clock64 trace,
    // profiles/r2_attn_fused_issuer_trace.txt) - and the softmax groups idled for exactly that long.
    {
      constexpr uint32_t idesc_g = make_idesc(uint32_t(FMT), 128, 96);
      int s = 0; uint32_t ph = 0;
      uint32_t xn_addr = 0;
      // G(gh): projection of global head gh into its slot's accumulator
      auto issue_g = [&](int gh) {
        const int h = gh % HEADS, ti = gh / HEADS, sl = gh % FA_NSLOT;
        const uint32_t n = uint32_t(gh / FA_NSLOT);
        const int xb = XNB == 2 ? (ti & 1) : 0;
        if (h == 0) {
          mbar_wait(&xn_full[xb], XNB == 2 ? ((ti >> 1) & 1) : (ti & 1));
          xn_addr = base + Cfg::XN_OFF + uint32_t(xb) * Cfg::XN_TILE;
        }
        mbar_wait(&acc_empty[sl], (n & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) FA_TR(0, gh);
        const uint32_t d_tmem = tmem_base + uint32_t(sl) * FA_TM_SLOT;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_full[s], ph);
          tc_fence_after();
          if (lane == 0 && kb < 2 && gh < 64) FA_TR(15, 2 * gh + kb);
          if (fa_elect_one()) {
            const uint64_t adesc = make_sw128_kmajor_desc(xn_addr + uint32_t(kb) * FA_SLAB);
            const uint64_t bdesc = make_sw128_kmajor_desc(base + Cfg::RING_OFF + uint32_t(s) * FA_STAGE);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc_g, (kb | k) ? 1u : 0u);
            umma_commit(&w_empty[s]);
            if (kb == KB - 1) {
              umma_commit(&acc_full[sl]);
              if (h == HEADS - 1) umma_commit(&xn_empty[xb]);     // the tile's last projection issued: xn may be overwritten once it completes
            }
          }
          __syncwarp();
          if (++s == FA_NST) { s = 0; ph ^= 1u; }
        }
        if (lane == 0) FA_TR(1, gh);
      };
      for (int gh = 0; gh < total; ++gh) issue_g(gh);
    }
  } else if (warp == 2) {
    // ---------------- attention issuer: per step s PV(s-2), then S(s) - the order in which their operands become ready in steady state
    {
      constexpr uint32_t idesc_s = make_idesc(uint32_t(FMT), 128, 64);
      constexpr uint32_t idesc_o = make_idesc(uint32_t(FMT), 128, 64) | (1u << 16);                // V' is MN-major
      for (int gs = 0; gs < total + 2; ++gs) {
        if (gs >= 2) {          // PV(gs - 2)
          const int gh = gs - 2, sl = gh % FA_NSLOT;
          mbar_wait(&p_full[sl], uint32_t(gh / FA_NSLOT) & 1u);
          tc_fence_after();
          if (lane == 0) FA_TR(2, gh);
          if (fa_elect_one()) {
            const uint32_t hb = base + Cfg::HB_OFF + uint32_t(sl) * FA_HB;
            const uint32_t d_tmem = tmem_base + uint32_t(sl) * FA_TM_SLOT + FA_TM_S;
            const uint64_t pdesc = make_sw128_kmajor_desc(hb + FA_QP);
            const uint64_t vdesc = fa_mnmajor_desc(hb + FA_V);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, pdesc + uint64_t(2 * k), vdesc + uint64_t(128 * k), idesc_o, k ? 1u : 0u);
            umma_commit(&o_full[sl]);
          }
          __syncwarp();
        }
        if (gs < total) {       // S(gs)
          const int sl = gs % FA_NSLOT;
          mbar_wait(&qkv_full[sl], uint32_t(gs / FA_NSLOT) & 1u);
          tc_fence_after();
          if (lane == 0) FA_TR(3, gs);
          if (fa_elect_one()) {
            const uint32_t hb = base + Cfg::HB_OFF + uint32_t(sl) * FA_HB;
            const uint32_t d_tmem = tmem_base + uint32_t(sl) * FA_TM_SLOT + FA_TM_S;
            const uint64_t qdesc = make_sw128_kmajor_desc(hb + FA_QP), kdesc = make_sw128_kmajor_desc(hb + FA_K);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss<false>(d_tmem, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
            umma_commit(&s_full[sl]);
          }
          __syncwarp();
          if (lane == 0) FA_TR(4, gs);
        }
      }
    }
  } else {
    // ---------------- spare warp 3: pull the rows of the tiles ahead into L2, at most three tiles ahead of the LayerNorm warps
    const int pl = lane;
    for (int ti = 1; ti < my_tiles; ++ti) {
      while (*ln_done + 3 < ti) __nanosleep(2000);      // (a tile takes ~6 us: polling faster only burns issue slots the softmax warps need)
      const int t = int(blockIdx.x) + ti * int(gridDim.x);
      for (int q = pl; q < FA_ROWS; q += 32) {
        const int wg = 2 * t + q / FA_L, i = q % FA_L;
        if (wg >= p.num_windows) continue;
        const int b = wg / p.nW, w = wg - b * p.nW;
        const char* src = reinterpret_cast<const char*>(p.x + (static_cast<long long>(b) * p.g.N + win_row_to_token(p.g, w * FA_L + i)) * C);
#pragma unroll
        for (int o = 0; o < C * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o));
      }
    }
  }
  } else if (warp < FA_LN_WARP0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------- softmax warpgroups: group g takes the global heads gh = g (mod 3) ----------------
    const int g = (warp - FA_SM_WARP0) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int wdx = r >> 6, j = r & 63;      // window of the pair, slot inside it
    const bool bf = FMT == 1;
    uint8_t* hb = smem + Cfg::HB_OFF + uint32_t(g) * FA_HB;
    const uint32_t lane_bits = uint32_t(quad * 32) << 16;
    const uint32_t ta = tmem_base + lane_bits + uint32_t(g) * FA_TM_SLOT;
    const uint32_t ts = ta + FA_TM_S;
    T16* ctx = static_cast<T16*>(p.ctx);
    long long tok_off = -1;
    uint32_t dm_lo = 0, dm_hi = 0;           // shift mask of this row: bit c set = key slot c lies in another region
    int cur_ti = -1;
    uint32_t n = 0;
    const int slot_iy = (j < FA_L ? j : 0) / 7, slot_ix = (j < FA_L ? j : 0) % 7;       // position of this row's slot inside its 7 x 7 window
    const int th = p.g.ws - p.g.shift, nWy = p.g.H / p.g.ws;
    const unsigned long long rows_lo = (1ull << (7 * th)) - 1ull;                     // key slots with iy < th
    const unsigned long long cols_lo = ((1ull << th) - 1ull) * 0x40810204081ull;      // key slots with ix < th
    for (int gh = g; gh < total; gh += FA_NSLOT, ++n) {
      const int ti = gh / HEADS, h = gh - ti * HEADS;
      const uint32_t ph = n & 1u;
      if (ti != cur_ti) {   // first head of a tile for this group: where the row goes, and its mask
        // (at C = 128 nearly every head of a group starts a new tile: the slot's (iy, ix) are hoisted out of the loop and the window's
        // (image, wy, wx) cost two divisions; win_row_to_token / fa_row_mask of common.cuh restated on those)
        cur_ti = ti;
        const int t = int(blockIdx.x) + ti * int(gridDim.x);
        const int wg = 2 * t + wdx;
        tok_off = -1; dm_lo = dm_hi = 0;
        if (j < FA_L && wg < p.num_windows) {
          const int b = wg / p.nW, w = wg - b * p.nW;
          const int wy = w / p.g.nWx, wx = w - wy * p.g.nWx;
          int y = wy * 7 + slot_iy + p.g.shift; if (y >= p.g.H) y -= p.g.H;
          int x = wx * 7 + slot_ix + p.g.shift; if (x >= p.g.W) x -= p.g.W;
          tok_off = (static_cast<long long>(b) * p.g.N + y * p.g.W + x) * C;
          if (p.g.shift > 0) {
            unsigned long long dm = 0;
            if (wy == nWy - 1) dm |= (slot_iy < th) ? ~rows_lo : rows_lo;
            if (wx == p.g.nWx - 1) dm |= (slot_ix < th) ? ~cols_lo : cols_lo;
            dm &= (1ull << FA_L) - 1ull;
            dm_lo = uint32_t(dm); dm_hi = uint32_t(dm >> 32);
          }
        }
      }
      // ---- D(h): projection accumulator -> Q' / K' / V'
      {
        if (quad == 0 && lane == 0) FA_TR(5, gh);
        mbar_wait(&acc_full[g], ph);
        tc_fence_after();
        if (quad == 0 && lane == 0) FA_TR(6, gh);
        // (the buffers of this slot are free: E(gh - 3) of this warp has seen PV(gh - 3) complete)
        uint8_t* qrow = hb + FA_QP + r * 128;
        uint8_t* krow = hb + FA_K + j * 128;
        uint8_t* vrow = hb + FA_V + j * 128;
        uint32_t rq[32], rk[32];
        tmem_ld_32x32(ta, rq);
        tmem_ld_wait();
        tmem_ld_32x32(ta + 32u, rk);
        {
          const float4* bq4 = reinterpret_cast<const float4*>(bq_tab + h * 32);
          const float2 qs2 = make_float2(p.qscale, p.qscale);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 b0 = bq4[2 * c], b1 = bq4[2 * c + 1];
            const float2 q0 = __ffma2_rn(make_float2(__uint_as_float(rq[8 * c]), __uint_as_float(rq[8 * c + 1])), qs2, make_float2(b0.x, b0.y));
            const float2 q1 = __ffma2_rn(make_float2(__uint_as_float(rq[8 * c + 2]), __uint_as_float(rq[8 * c + 3])), qs2, make_float2(b0.z, b0.w));
            const float2 q2 = __ffma2_rn(make_float2(__uint_as_float(rq[8 * c + 4]), __uint_as_float(rq[8 * c + 5])), qs2, make_float2(b1.x, b1.y));
            const float2 q3 = __ffma2_rn(make_float2(__uint_as_float(rq[8 * c + 6]), __uint_as_float(rq[8 * c + 7])), qs2, make_float2(b1.z, b1.w));
            const uint32_t own = uint32_t(((wdx * 4 + c) ^ (r & 7)) << 4), other = uint32_t((((1 - wdx) * 4 + c) ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(qrow + own) = make_uint4(pack16(bf, q0.x, q0.y), pack16(bf, q1.x, q1.y), pack16(bf, q2.x, q2.y), pack16(bf, q3.x, q3.y));
            *reinterpret_cast<uint4*>(qrow + other) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        auto store_kv = [&](uint8_t* row, const uint32_t (&rr)[32]) {
          if (j < FA_L) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(row + (((wdx * 4 + c) ^ (r & 7)) << 4)) =
                  make_uint4(pack16(bf, __uint_as_float(rr[8 * c]), __uint_as_float(rr[8 * c + 1])), pack16(bf, __uint_as_float(rr[8 * c + 2]), __uint_as_float(rr[8 * c + 3])),
                             pack16(bf, __uint_as_float(rr[8 * c + 4]), __uint_as_float(rr[8 * c + 5])), pack16(bf, __uint_as_float(rr[8 * c + 6]), __uint_as_float(rr[8 * c + 7])));
          }
        };
        tmem_ld_wait();
        tmem_ld_32x32(ta + 64u, rq);      // v, into the registers q has left
        store_kv(krow, rk);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[g]);   // accumulator drained: the projection of head gh + 3 may overwrite it
        store_kv(vrow, rq);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qkv_full[g]);
        if (quad == 0 && lane == 0) FA_TR(7, gh);
      }
      // ---- X(h): logits -> unnormalised probabilities (log2 domain), written over Q'
      float rsum;
      {
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        if (quad == 0 && lane == 0) FA_TR(8, gh);
        uint32_t s0[32], s1[16], s2;
        tmem_ld_32x32(ts, s0);
        tmem_ld_32x16(ts + 32u, s1);
        tmem_ld_32x1(ts + 48u, s2);
        tmem_ld_wait();
        tc_fence_before();
        float sv[50];
#pragma unroll
        for (int c = 0; c < 32; ++c) sv[c] = __uint_as_float(s0[c]);
#pragma unroll
        for (int c = 0; c < 16; ++c) sv[32 + c] = __uint_as_float(s1[c]);
        sv[48] = __uint_as_float(s2);
        sv[49] = 0.0f;
        {   // + relative-position bias of this query slot (row jb of the head's table; padding rows read a real row, their output is dropped)
          const int bs = gh % FA_NBS;
          mbar_wait(&b_full[bs], (uint32_t(gh) / FA_NBS) & 1u);
          const uint4* brow = reinterpret_cast<const uint4*>(smem + Cfg::BIAS_OFF + uint32_t(bs) * FA_BIAS_STAGE + (j < FA_L ? j : 0) * 112);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            const uint4 b4 = brow[c];
            fa_add_h2(sv[8 * c], sv[8 * c + 1], b4.x);
            if (c < 6) {
              fa_add_h2(sv[8 * c + 2], sv[8 * c + 3], b4.y);
              fa_add_h2(sv[8 * c + 4], sv[8 * c + 5], b4.z);
              fa_add_h2(sv[8 * c + 6], sv[8 * c + 7], b4.w);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_empty[bs]);
        }
        sv[49] = -INFINITY;
        fa_add_mask(sv, dm_lo, dm_hi, p.g.ws - p.g.shift == 4);
        float mx0 = fa_max3(sv[0], sv[1], sv[2]), mx1 = fa_max3(sv[3], sv[4], sv[5]);
#pragma unroll
        for (int c = 6; c < 48; c += 4) { mx0 = fa_max3(mx0, sv[c], sv[c + 1]); mx1 = fa_max3(mx1, sv[c + 2], sv[c + 3]); }
        const float mxx = fmaxf(mx0, mx1);               // (the loop ends with c = 46: sv[46..49], sv[49] = -inf)
        const float2 nm = make_float2(-mxx, -mxx);
        float2 acc2 = make_float2(0.f, 0.f);
        uint32_t pp[25];
#pragma unroll
        for (int c = 0; c < 25; ++c) {
          const float2 d = __fadd2_rn(make_float2(sv[2 * c], sv[2 * c + 1]), nm);
          const float2 e = make_float2(fa_exp2(d.x), c == 24 ? 0.0f : fa_exp2(d.y));
          acc2 = __fadd2_rn(acc2, e);
          pp[c] = pack16(bf, e.x, e.y);
        }
        rsum = acc2.x + acc2.y;
        uint8_t* prow = hb + FA_QP + r * 128;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        *reinterpret_cast<uint4*>(prow + ((6 ^ (r & 7)) << 4)) = make_uint4(pp[24], 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(prow + ((7 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        if (quad == 0 && lane == 0) FA_TR(9, gh);
      }
      // ---- E(h): O' / rowsum + bv -> token-ordered context
      {
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        if (quad == 0 && lane == 0) FA_TR(10, gh);
        uint32_t o[32];
        tmem_ld_32x32(ts + uint32_t(wdx * 32), o);
        tmem_ld_wait();
        tc_fence_before();     // S(gh + 3) overwrites these columns only after this group's next qkv_full arrive
        if (tok_off >= 0) {
          const float inv = 1.0f / rsum;
          const float2 inv2 = make_float2(inv, inv);
          const float4* bv4 = reinterpret_cast<const float4*>(bv_tab + h * 32);
          uint4* dst = reinterpret_cast<uint4*>(ctx + tok_off + h * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = bv4[2 * q], b1 = bv4[2 * q + 1];
            const float2 c0 = __ffma2_rn(make_float2(__uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])), inv2, make_float2(b0.x, b0.y));
            const float2 c1 = __ffma2_rn(make_float2(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])), inv2, make_float2(b0.z, b0.w));
            const float2 c2 = __ffma2_rn(make_float2(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])), inv2, make_float2(b1.x, b1.y));
            const float2 c3 = __ffma2_rn(make_float2(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])), inv2, make_float2(b1.z, b1.w));
            dst[q] = make_uint4(pack16(bf, c0.x, c0.y), pack16(bf, c1.x, c1.y), pack16(bf, c2.x, c2.y), pack16(bf, c3.x, c3.y));
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    // ---------------- LayerNorm producers: gathered fp32 rows -> normalised 16-bit tile in the MMA layout ----------------
    // A lane owns two 8-column chunks of a row (C / 16 lanes per row); a warp pass covers 32 / LPR rows.  Each of the 8 warps keeps
    // two passes in registers, normalised together (their shuffle chains interleave) and requested again at once: the load latency
    // of one warp is covered by the arithmetic of the other seven.
    constexpr int LPR = C / 16;                       // lanes per row
    constexpr int RPP = 32 / LPR;                     // rows per warp pass
    constexpr int CPL = 2;                            // 8-column chunks per lane
    constexpr int NPASS = (FA_ROWS + FA_LN_WARPS * RPP - 1) / (FA_LN_WARPS * RPP);
    constexpr int GP = 2, GH = 2;                     // passes in registers, passes per group
    constexpr int NB = (NPASS + GP - 1) / GP;         // batches per tile
    const int ln = warp - FA_LN_WARP0;
    const int sub = lane / LPR, lr = lane % LPR;
    int* rowtab = reinterpret_cast<int*>(smem + Cfg::HB_OFF + 2 * FA_HB + FA_TAB);   // [3][128] token row of tile row q (-1: none), slot 2's K' padding
    // one thread per tile row works out where the row lives (closed-form shift / partition map), once per tile, one tile ahead
    auto write_table = [&](int ti) {
      if (ln >= 4) return;
      const int q = ln * 32 + lane;
      int tok = -1;
      if (ti < my_tiles && q < FA_ROWS) {
        const int wg = 2 * (int(blockIdx.x) + ti * int(gridDim.x)) + q / FA_L, i = q % FA_L;
        if (wg < p.num_windows) {
          const int b = wg / p.nW, w = wg - b * p.nW;
          tok = b * p.g.N + win_row_to_token(p.g, w * FA_L + i);
        }
      }
      rowtab[(ti % 3) * 128 + q] = tok;
    };
    // request the raw row of pass `pidx` of tile `ti` (rows that do not exist read token 0 and are dropped at the store)
    auto issue = [&](int ti, int pidx, float2 (&vv)[CPL][4], int& rr) {
      const int q = (pidx * FA_LN_WARPS + ln) * RPP + sub;
      int tok = -1;
      if (ti < my_tiles && pidx < NPASS && q < FA_ROWS) tok = rowtab[(ti % 3) * 128 + q];
      rr = tok >= 0 ? (q / FA_L) * 64 + q % FA_L : -1;
      const float* xr = p.x + static_cast<long long>(tok < 0 ? 0 : tok) * C;
#pragma unroll
      for (int tt = 0; tt < CPL; ++tt) {
        const int c = lr + LPR * tt;
        const float4 a0 = *reinterpret_cast<const float4*>(xr + c * 8), a1 = *reinterpret_cast<const float4*>(xr + c * 8 + 4);
        vv[tt][0] = make_float2(a0.x, a0.y); vv[tt][1] = make_float2(a0.z, a0.w);
        vv[tt][2] = make_float2(a1.x, a1.y); vv[tt][3] = make_float2(a1.z, a1.w);
      }
    };
    float2 v[GP][CPL][4];
    int rowq[GP];
    write_table(0);
    write_table(1);
    asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
    for (int pp = 0; pp < GP; ++pp) issue(0, pp, v[pp], rowq[pp]);
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int xb = XNB == 2 ? (ti & 1) : 0;
      uint8_t* xn = smem + Cfg::XN_OFF + uint32_t(xb) * Cfg::XN_TILE;
      if (ln == 0 && lane == 0) FA_TR(11, ti);
      if (ti > 0) {
        write_table(ti + 1);    // (buffer of tile ti - 2: every warp has passed the barrier of tile ti - 1, i.e. finished tile ti - 2)
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (ln == 0 && lane == 0) FA_TR(12, ti);
      mbar_wait(&xn_empty[xb], (XNB == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1u);
      if (ln == 0 && lane == 0) FA_TR(13, ti);
#pragma unroll 1
      for (int bt = 0; bt < NB; ++bt) {
        const bool last = bt + 1 == NB;
        const int nti = last ? ti + 1 : ti, np0 = last ? 0 : (bt + 1) * GP;
#pragma unroll
        for (int g0 = 0; g0 < GP; g0 += GH) {
          if (bt * GP + g0 < NPASS) {      // (warp-uniform) groups past the last pass hold no rows
            float mean[GH], rstd[GH];
#pragma unroll
            for (int u = 0; u < GH; ++u) {
              float2 (&w)[CPL][4] = v[g0 + u];
              const float2 s2 = __fadd2_rn(__fadd2_rn(__fadd2_rn(w[0][0], w[0][1]), __fadd2_rn(w[0][2], w[0][3])),
                                           __fadd2_rn(__fadd2_rn(w[1][0], w[1][1]), __fadd2_rn(w[1][2], w[1][3])));
              mean[u] = s2.x + s2.y;
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
              for (int u = 0; u < GH; ++u) mean[u] += __shfl_xor_sync(0xffffffffu, mean[u], o);
#pragma unroll
            for (int u = 0; u < GH; ++u) {
              float2 (&w)[CPL][4] = v[g0 + u];
              mean[u] *= 1.0f / float(C);
              const float2 nm = make_float2(-mean[u], -mean[u]);
              float2 qa = make_float2(0.f, 0.f), qb = make_float2(0.f, 0.f);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                w[0][e] = __fadd2_rn(w[0][e], nm); qa = __ffma2_rn(w[0][e], w[0][e], qa);
                w[1][e] = __fadd2_rn(w[1][e], nm); qb = __ffma2_rn(w[1][e], w[1][e], qb);
              }
              rstd[u] = (qa.x + qa.y) + (qb.x + qb.y);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
              for (int u = 0; u < GH; ++u) rstd[u] += __shfl_xor_sync(0xffffffffu, rstd[u], o);
#pragma unroll
            for (int u = 0; u < GH; ++u) {
              float2 (&w)[CPL][4] = v[g0 + u];
              const float rs = rsqrtf(rstd[u] * (1.0f / float(C)) + p.eps);
              const float2 rs2 = make_float2(rs, rs);
              const int rr = rowq[g0 + u];
              const int ra = rr < 0 ? 0 : rr;
#pragma unroll
              for (int tt = 0; tt < CPL; ++tt) {
                const int c = lr + LPR * tt;
                float2 y[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fmul2_rn(w[tt][e], rs2);     // (gamma / beta live in the packed weights and biases)
                uint4 pk;
                pk.x = Half16<T16>::pack(y[0].x, y[0].y); pk.y = Half16<T16>::pack(y[1].x, y[1].y);
                pk.z = Half16<T16>::pack(y[2].x, y[2].y); pk.w = Half16<T16>::pack(y[3].x, y[3].y);
                if (rr >= 0) *reinterpret_cast<uint4*>(xn + (c >> 3) * FA_SLAB + ra * 128 + (((c & 7) ^ (ra & 7)) << 4)) = pk;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < GH; ++u) issue(nti, np0 + g0 + u, v[g0 + u], rowq[g0 + u]);   // the registers are free: request what they hold next
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (ln == 0 && lane == 0) FA_TR(14, ti);
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
static int launch_fa(const CUtensorMap& tmW, const FaParams& p, cudaStream_t stream) {
  using Cfg = FaCfg<C>;
  auto kern = swin_attn_fused_kernel<FMT, C>;
  CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM)));
  const int tiles = (p.num_windows + 1) / 2;
  const int ctas = tiles < num_sms() ? tiles : num_sms();
  CSVIT_CUDA(launch_pdl(kern, dim3(ctas), dim3(FA_THREADS), Cfg::SMEM, stream, tmW, p));
  return 0;
}

int launch_swin_attn_fused(const float* x, float eps, const void* wqkv_h, const float* bqkv_h,
                           const void* bias_op, void* ctx, int dtype, int B, int H, int W, int C, int heads, int ws, int shift,
                           cudaStream_t stream) {
  CSVIT_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "swin_attn_fused: 16-bit operand formats only");
  CSVIT_REQUIRE(C == 128 || C == 256, "swin_attn_fused: C=%d not in {128, 256}", C);
  CSVIT_REQUIRE(ws == 7 && C == heads * 32, "swin_attn_fused: window 7 / head_dim 32 only (ws=%d C=%d heads=%d)", ws, C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "swin_attn_fused: bad geometry %dx%d shift %d", H, W, shift);
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bqkv_h) & 15) == 0,
                "swin_attn_fused: operands must be 16-byte aligned");
  const int nW = (H / ws) * (W / ws);
  const long long windows = static_cast<long long>(B) * nW;
  if (windows <= 0) return 0;
  CSVIT_REQUIRE(windows < (1ll << 30), "swin_attn_fused: too many windows");
  FaParams p{};
  p.x = x; p.eps = eps; p.bqkv = bqkv_h; p.bias = bias_op; p.ctx = ctx;
  p.num_windows = static_cast<int>(windows); p.nW = nW;
  p.qscale = 1.4426950408889634f * 0.17677669529663687f;
  p.g = make_geom(H, W, ws, shift);
  CUtensorMap tmW;
  if (int e = make_tmap(&tmW, wqkv_h, C, 3ll * C, C, dtype, 96, true)) return e;
  const bool bf = dtype == DT_BF16;
  if (C == 128) return bf ? launch_fa<1, 128>(tmW, p, stream) : launch_fa<0, 128>(tmW, p, stream);
  return bf ? launch_fa<1, 256>(tmW, p, stream) : launch_fa<0, 256>(tmW, p, stream);
}

}  // namespace csvit

#ifdef CSVIT_FA_TRACE
extern "C" __attribute__((visibility("default"))) int csvit_debug_fa_trace(long long* host_out) {
  return int(cudaMemcpyFromSymbol(host_out, csvit::fa_trace_buf, sizeof(long long) * 16 * 256));
}
#endif
