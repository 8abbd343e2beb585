// Shifted-window attention core on tcgen05 / TMEM for every Swin width (C = 32 * heads): from the window-ordered 16-bit
// [rows, 3C] output of the Q/K/V GEMM to the attention context,
//     ctx = concat_h softmax(q_h k_h^T / sqrt(32) + bias_h + shift_mask) v_h
// i.e. transpose_for_scores + Q K^T + relative-position-bias gather + mask add + softmax + P V + head merge (+ window_reverse and
// roll(+s) when the context is written in token order) of HF:swin/modeling_swin.py:404-459, 556-582, 631-636.
// It replaces the round-1 mma.sync kernel: operands come in by TMA, both contractions run on tcgen05 with TMEM accumulators.
//
// Tile = two consecutive 49-token windows (98 consecutive rows of qkv), one HEAD PAIR (64 columns of q, of k and of v) per
// pipeline stage.  TMA boxes of {64 columns, 49 rows} with the 128-byte swizzle are exactly the operand tiles tcgen05 reads:
//   Q, K   [128 rows x 128 B] K-major: window A in rows 0-48, window B in rows 64-112 (two boxes each); head e of the pair is the
//          k-slice of columns 32e .. 32e+31, i.e. descriptor start + 64 e bytes - no register pass, no repacking
//   V'     [2 chunks][64 keys x 128 B] MN-major: chunk 0 = window A's values (both heads), chunk 1 = window B's
//   S(e)   S[128 x 128] = Q_e K_e^T  (2 k-steps, N = 128): row r of window w(r) finds its logits in columns 64 w(r) .. +48; the
//          cross-window quadrants are computed and ignored (an M = 128 MMA step costs the same for every N <= 128)
//   X(e)   one thread per row: 49 logits from TMEM + fp16 log2-domain bias row + closed-form shift mask, exp2, unnormalised P
//          (16 bit) written compactly [128 rows x 64 keys] over the pair's Q tile (e = 0) or K tile (e = 1) - dead once both S are done
//   PV(e)  O'[128 x 128] = P_e[128 x 64] V'[64 x 128] (4 k-steps, over S in TMEM): row r reads columns 64 w(r) + 32 e .. +31
//   E(e)   O' / rowsum -> 16 bit -> ctx row (token order: window_reverse + un-shift folded into the store address)
// Roles: warp 0 TMA producer (4-stage ring of head pairs, 48 KB each), warp 2 bias-table producer (4-deep ring), warp 1 MMA issuer,
// warps 4-19 four softmax warpgroups: global head gh uses TMEM slot gh % 4 (128 columns) and warpgroup gh % 4, continuously
// across tile boundaries, so four heads are in flight and the tensor pipe, the MUFU/ALU work and the loads overlap.
// Per token the kernel reads 6C bytes and writes 2C: it is HBM-bound (8C B/token) from C = 128 to 1024.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "attn_common.cuh"
#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"

namespace csvit {

constexpr int AC_THREADS = 640;                 // 5 warpgroups: {TMA, MMA, 2 idle}, 4 x softmax
constexpr int AC_NST = 4;                       // head-pair stages
constexpr int AC_NSLOT = 4;                     // TMEM slots / softmax groups / bias stages
constexpr uint32_t AC_TILE = 128 * 128;         // 128 rows x 64 16-bit columns
constexpr uint32_t AC_STAGE = 3 * AC_TILE;      // Q | K | V'
constexpr uint32_t AC_BOX = 49 * 128;           // bytes of one TMA box
constexpr uint32_t AC_BIAS_OFF = AC_NST * AC_STAGE;
constexpr uint32_t AC_BAR_OFF = AC_BIAS_OFF + AC_NSLOT * FA_BIAS_STAGE;
constexpr size_t AC_SMEM = 1024 + size_t(AC_BAR_OFF) + 512;

struct AcParams {
  const void* bias;      // fp16 [heads][49][56]: log2(e) * relative position bias of (query slot, key slot)
  void* ctx;             // 16-bit [B*N, C]
  int num_windows;       // B * nW
  int nW;
  int C, heads;
  int token_order;       // 1: ctx rows in token order (window_reverse + roll(+s) folded in); 0: window order like qkv
  float qscale;          // log2(e) / sqrt(32), or 1 when the Q/K/V GEMM already applied it to q
  WinGeom g;
  long long* trace;      // debugging (build with -DCSVIT_AC_TRACE_BUILD, run with CSVIT_AC_TRACE=<file>): clock64 stamps of CTA 0
};
#ifdef CSVIT_AC_TRACE_BUILD
constexpr int AC_TRACE_HEADS = 24;
#define AC_STAMP(k) do { if (tr) tr[k] = clock64(); } while (0)
#else
#define AC_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ uint64_t ac_mnmajor_desc(uint32_t smem_addr) {
  // MN-major SWIZZLE_128B operand, N = 128 = two 128-byte chunks: LBO = chunk stride (64 key rows), SBO = 8 key rows
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// FMT: 0 = fp16, 1 = bf16.  SCALED: logits still need the 1/sqrt(32) (and log2 e) factor.
template <int FMT, bool SCALED>
__global__ void __launch_bounds__(AC_THREADS, 1)
swin_attn_core_kernel(const __grid_constant__ CUtensorMap tmQ, AcParams p) {
  using T16 = typename std::conditional<FMT == 1, __nv_bfloat16, __half>::type;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AC_BAR_OFF);
  uint64_t* st_full = bars;                    // [AC_NST]
  uint64_t* st_empty = st_full + AC_NST;       // [AC_NST]
  uint64_t* s_full = st_empty + AC_NST;        // [AC_NSLOT] by TMEM slot, like everything below
  uint64_t* p_full = s_full + AC_NSLOT;
  uint64_t* o_full = p_full + AC_NSLOT;
  uint64_t* o_empty = o_full + AC_NSLOT;
  uint64_t* b_full = o_empty + AC_NSLOT;
  uint64_t* b_empty = b_full + AC_NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + AC_NSLOT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, HEADS = p.heads;
  const int PAIRS = (HEADS + 1) >> 1;           // an odd head count (Swin-T: 3) leaves a phantom head that is computed and dropped
  // work unit = (tile, head pair); every CTA takes one contiguous range of units (tile-major), so the grid is balanced to one
  // unit (Swin-B stage 2 at batch 256: 4096 units over 148 CTAs instead of 512 tiles) and a CTA walks adjacent rows of qkv
  const int num_tiles = (p.num_windows + 1) >> 1;
  const long long units = static_cast<long long>(num_tiles) * PAIRS;
  const int u_begin = int(units * blockIdx.x / gridDim.x), u_end = int(units * (blockIdx.x + 1) / gridDim.x);
  const int total_pairs = u_end - u_begin;

  // padding rows (49-63, 113-127 of Q / K, 49-63 of the V' chunks) are never written by TMA: they must hold finite values
  for (uint32_t i = threadIdx.x; i < AC_BIAS_OFF / 16; i += AC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    for (int s = 0; s < AC_NST; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1); }
    for (int b = 0; b < AC_NSLOT; ++b) {
      mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 4);
      mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 4);
      mbar_init(&b_full[b], 1); mbar_init(&b_empty[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();      // PDL: the next kernel's prologue may overlap this kernel's tail ...
  griddep_wait();        // ... and this kernel touches global memory only after its predecessors have completed

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer: one head pair per stage (6 boxes), one bias table per head
      for (int gp = 0; gp < total_pairs; ++gp) {
        const int st = gp % AC_NST;
        const int t = (u_begin + gp) / PAIRS, hp = (u_begin + gp) - t * PAIRS;
        const int row0 = t * (2 * FA_L);
        mbar_wait(&st_empty[st], ((uint32_t(gp) / AC_NST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&st_full[st], 6 * AC_BOX);
        uint8_t* sb = smem + size_t(st) * AC_STAGE;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const int col = m * C + hp * 64;
          // Q / K: window B starts at tile row 64; V': window B is the second MN chunk (64 key rows further)
          tma_load_2d(sb + m * AC_TILE, &tmQ, &st_full[st], col, row0);
          tma_load_2d(sb + m * AC_TILE + 8192, &tmQ, &st_full[st], col, row0 + FA_L);
        }
      }
    } else if (warp == 2 && lane == 0) {
      // ---------------- bias producer: one table per head on its own ring (L2-resident, short latency), so that the operand
      // ring above runs its full depth ahead instead of being paced by the softmax groups' bias releases
      const int total_heads = 2 * total_pairs;
      for (int gh = 0; gh < total_heads; ++gh) {
        const int bs = gh & (AC_NSLOT - 1);
        const int h = min((2 * u_begin + gh) % (2 * PAIRS), HEADS - 1);
        mbar_wait(&b_empty[bs], ((uint32_t(gh) / AC_NSLOT) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&b_full[bs], FA_BIAS_BYTES);
        fa_bulk_load(smem + AC_BIAS_OFF + uint32_t(bs) * FA_BIAS_STAGE, static_cast<const char*>(p.bias) + size_t(h) * FA_BIAS_BYTES,
                     FA_BIAS_BYTES, &b_full[bs]);
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer: per step S(gp) for both heads, then PV(gp - 1) for both heads (converged warp, one elected lane
      // issues: see fa_elect_one)
      constexpr uint32_t idesc_s = make_idesc(uint32_t(FMT), 128, 128);
      constexpr uint32_t idesc_o = make_idesc(uint32_t(FMT), 128, 128) | (1u << 16);      // V' is MN-major
      for (int gp = 0; gp <= total_pairs; ++gp) {
        if (gp < total_pairs) {
          const int st = gp % AC_NST;
          mbar_wait(&st_full[st], (uint32_t(gp) / AC_NST) & 1u);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t gh = uint32_t(2 * gp + e);
            mbar_wait(&o_empty[gh & 3u], ((gh >> 2) & 1u) ^ 1u);       // E(gh - 4) has drained this slot's columns
          }
          tc_fence_after();
          if (fa_elect_one()) {
            const uint32_t sb = base + uint32_t(st) * AC_STAGE;
            const uint64_t qdesc = make_sw128_kmajor_desc(sb), kdesc = make_sw128_kmajor_desc(sb + AC_TILE);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const uint32_t d_tmem = tmem_base + (uint32_t(2 * gp + e) & 3u) * 128u;
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_ss<false>(d_tmem, qdesc + uint64_t(2 * (2 * e + k)), kdesc + uint64_t(2 * (2 * e + k)), idesc_s, k ? 1u : 0u);
            }
            // both arrive once BOTH products are complete: the probabilities overwrite the Q and K tiles
            umma_commit(&s_full[uint32_t(2 * gp) & 3u]);
            umma_commit(&s_full[uint32_t(2 * gp + 1) & 3u]);
          }
          __syncwarp();
        }
        if (gp >= 1) {
          const int q = gp - 1, st = q % AC_NST;
          const uint32_t sb = base + uint32_t(st) * AC_STAGE;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t gh = uint32_t(2 * q + e), sl = gh & 3u;
            mbar_wait(&p_full[sl], (gh >> 2) & 1u);
            tc_fence_after();
            if (fa_elect_one()) {
              const uint64_t vdesc = ac_mnmajor_desc(sb + 2 * AC_TILE);
              const uint64_t pdesc = make_sw128_kmajor_desc(sb + uint32_t(e) * AC_TILE);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss<false>(tmem_base + sl * 128u, pdesc + uint64_t(2 * k), vdesc + uint64_t(128 * k), idesc_o, k ? 1u : 0u);
              umma_commit(&o_full[sl]);
              if (e == 1) umma_commit(&st_empty[st]);         // the pair's stage may be reloaded once both P V products are complete
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // (the register pool is what the CTA was launched with: 640 x 96; warps 0-3 hand back 128 x 56, enough for 512 x 8 more)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------- softmax warpgroups: group g takes the global heads gh = g (mod 4) ----------------
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int wdx = r >> 6, j = r & 63;      // window of the pair, slot inside it
    const bool bf = FMT == 1;
    const uint32_t ts = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(g) * 128u + uint32_t(wdx) * 64u;
    T16* ctx = static_cast<T16*>(p.ctx);
    long long tok_off = -1;
    uint32_t dm_lo = 0, dm_hi = 0;           // shift mask of this row: bit c set = key slot c lies in another region
    int cur_ti = -1;
    const int total_heads = 2 * total_pairs;
    uint32_t n = 0;
    // (tile, head pair) of this group's current head, advanced without divisions: the group's heads are AC_NSLOT = 4 apart = 2 pairs
    int ti = (u_begin + (g >> 1)) / PAIRS, hp = (u_begin + (g >> 1)) - ti * PAIRS;
    const int ti_step = 2 / PAIRS, hp_step = 2 % PAIRS;
    for (int gh = g; gh < total_heads; gh += AC_NSLOT, ++n) {
      const int gp = gh >> 1, e = gh & 1;
      const int h = 2 * hp + e;
      const uint32_t ph = n & 1u;
      uint8_t* sb = smem + size_t(gp % AC_NST) * AC_STAGE;
#ifdef CSVIT_AC_TRACE_BUILD
      long long* tr = (p.trace && blockIdx.x == 0 && quad == 0 && lane == 0 && n < AC_TRACE_HEADS) ? p.trace + (g * AC_TRACE_HEADS + n) * 8 : nullptr;
#endif
      AC_STAMP(0);
      if (ti != cur_ti) {   // first head of a tile for this group: where the row goes, and its mask
        cur_ti = ti;
        const int wg = 2 * ti + wdx;
        tok_off = -1; dm_lo = dm_hi = 0;
        if (j < FA_L && wg < p.num_windows) {
          const int b = wg / p.nW, w = wg - b * p.nW;
          const long long row = p.token_order ? static_cast<long long>(b) * p.g.N + win_row_to_token(p.g, w * FA_L + j)
                                              : static_cast<long long>(wg) * FA_L + j;
          tok_off = row * C;
          const unsigned long long dm = fa_row_mask(p.g, w, j);
          dm_lo = uint32_t(dm); dm_hi = uint32_t(dm >> 32);
        }
      }
      // ---- X: logits -> unnormalised probabilities (log2 domain), written over the pair's Q (e = 0) or K (e = 1) tile
      float rsum;
      {
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        AC_STAMP(1);
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
        {   // + relative-position bias of this query slot (padding rows read a real row, their output is dropped)
          AC_STAMP(2);
          mbar_wait(&b_full[g], ph);
          AC_STAMP(3);
          const uint4* brow = reinterpret_cast<const uint4*>(smem + AC_BIAS_OFF + uint32_t(g) * FA_BIAS_STAGE + (j < FA_L ? j : 0) * 112);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            const uint4 b4 = brow[c];
            if constexpr (SCALED) {
              const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&b4.x));
              sv[8 * c] = fmaf(sv[8 * c], p.qscale, f0.x); sv[8 * c + 1] = fmaf(sv[8 * c + 1], p.qscale, f0.y);
              if (c < 6) {
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&b4.y));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&b4.z));
                const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&b4.w));
                sv[8 * c + 2] = fmaf(sv[8 * c + 2], p.qscale, f1.x); sv[8 * c + 3] = fmaf(sv[8 * c + 3], p.qscale, f1.y);
                sv[8 * c + 4] = fmaf(sv[8 * c + 4], p.qscale, f2.x); sv[8 * c + 5] = fmaf(sv[8 * c + 5], p.qscale, f2.y);
                sv[8 * c + 6] = fmaf(sv[8 * c + 6], p.qscale, f3.x); sv[8 * c + 7] = fmaf(sv[8 * c + 7], p.qscale, f3.y);
              }
            } else {
              fa_add_h2(sv[8 * c], sv[8 * c + 1], b4.x);
              if (c < 6) {
                fa_add_h2(sv[8 * c + 2], sv[8 * c + 3], b4.y);
                fa_add_h2(sv[8 * c + 4], sv[8 * c + 5], b4.z);
                fa_add_h2(sv[8 * c + 6], sv[8 * c + 7], b4.w);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_empty[g]);
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
          const float2 ex = make_float2(fa_exp2(d.x), c == 24 ? 0.0f : fa_exp2(d.y));
          acc2 = __fadd2_rn(acc2, ex);
          pp[c] = pack16(bf, ex.x, ex.y);
        }
        rsum = acc2.x + acc2.y;
        uint8_t* prow = sb + uint32_t(e) * AC_TILE + r * 128;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]);
        *reinterpret_cast<uint4*>(prow + ((6 ^ (r & 7)) << 4)) = make_uint4(pp[24], 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(prow + ((7 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        AC_STAMP(4);
      }
      // ---- E: O' / rowsum -> context row
      {
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        AC_STAMP(5);
        uint32_t o[32];
        tmem_ld_32x32(ts + uint32_t(e) * 32u, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[g]);      // S(gh + 4) may overwrite the slot
        if (tok_off >= 0 && h < HEADS) {
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
      AC_STAMP(6);
      ti += ti_step; hp += hp_step;
      if (hp >= PAIRS) { hp -= PAIRS; ++ti; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int FMT, bool SCALED>
static int launch_ac(const CUtensorMap& tmQ, const AcParams& p, cudaStream_t stream) {
  auto kern = swin_attn_core_kernel<FMT, SCALED>;
  CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AC_SMEM)));
  const long long units = static_cast<long long>((p.num_windows + 1) / 2) * ((p.heads + 1) / 2);
  const int ctas = units < num_sms() ? int(units) : num_sms();
  CSVIT_CUDA(launch_pdl(kern, dim3(ctas), dim3(AC_THREADS), AC_SMEM, stream, tmQ, p));
  return 0;
}

// qkv: 16-bit [B*nW*49, 3C] rows in window order (row pitch ldq elements), per row q | k | v with head h in columns 32h .. 32h+31.
int launch_swin_attn_core(const void* qkv, long long ldq, const void* bias_log2, void* ctx, int dtype, int B, int H, int W, int C,
                          int heads, int ws, int shift, int token_order, int q_prescaled, cudaStream_t stream) {
  CSVIT_REQUIRE(dtype == DT_BF16 || dtype == DT_F16, "swin_attn_core: 16-bit operand formats only");
  CSVIT_REQUIRE(ws == 7 && C == heads * 32 && heads >= 1, "swin_attn_core: window 7 / head_dim 32 only (ws=%d C=%d heads=%d)", ws, C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0 && shift >= 0 && shift < ws, "swin_attn_core: bad geometry %dx%d shift %d", H, W, shift);
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias_log2) & 15) == 0 && ldq >= 3ll * C,
                "swin_attn_core: operands must be 16-byte aligned, pitch >= 3C");
  const int nW = (H / ws) * (W / ws);
  const long long windows = static_cast<long long>(B) * nW;
  if (windows <= 0) return 0;
  CSVIT_REQUIRE(windows * FA_L < (1ll << 31), "swin_attn_core: too many rows");
  AcParams p{};
  p.trace = nullptr;
  p.bias = bias_log2; p.ctx = ctx;
  p.num_windows = static_cast<int>(windows); p.nW = nW; p.C = C; p.heads = heads;
  p.token_order = token_order ? 1 : 0;
  p.qscale = q_prescaled ? 1.0f : 1.4426950408889634f * 0.17677669529663687f;
  p.g = make_geom(H, W, ws, shift);
  CUtensorMap tmQ;
  if (int e = make_tmap(&tmQ, qkv, ldq, windows * FA_L, 3ll * C, dtype, FA_L, false)) return e;
  const bool bf = dtype == DT_BF16;
#ifdef CSVIT_AC_TRACE_BUILD
  if (const char* path = getenv("CSVIT_AC_TRACE")) {
    const size_t nb = size_t(AC_NSLOT) * AC_TRACE_HEADS * 8 * sizeof(long long);
    CSVIT_CUDA(cudaMalloc(&p.trace, nb));
    CSVIT_CUDA(cudaMemsetAsync(p.trace, 0, nb, stream));
    if (int e = bf ? launch_ac<1, false>(tmQ, p, stream) : launch_ac<0, false>(tmQ, p, stream)) return e;
    CSVIT_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> h(nb / sizeof(long long));
    CSVIT_CUDA(cudaMemcpy(h.data(), p.trace, nb, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(path, "w")) {
      long long t0 = 0;
      for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
      fprintf(f, "# swin_attn_core, CTA 0, one warp per softmax group; per head: loop top | S wait | ld + prep | bias wait | softmax + P store | PV wait | E\n");
      for (int g = 0; g < AC_NSLOT; ++g)
        for (int n = 0; n < AC_TRACE_HEADS; ++n) {
          const long long* r = h.data() + (g * AC_TRACE_HEADS + n) * 8;
          if (!r[0]) continue;
          fprintf(f, "g%d n%2d start %8lld | %6lld %6lld %6lld %6lld %6lld %6lld | total %6lld\n", g, n, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2],
                  r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6] - r[0]);
        }
      fclose(f);
    }
    return 0;
  }
#endif
  if (q_prescaled) return bf ? launch_ac<1, false>(tmQ, p, stream) : launch_ac<0, false>(tmQ, p, stream);
  return bf ? launch_ac<1, true>(tmQ, p, stream) : launch_ac<0, true>(tmQ, p, stream);
}

}  // namespace csvit
