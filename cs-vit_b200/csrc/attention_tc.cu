// Swin window attention core on tcgen05 / TMEM: two 49-token windows packed block-diagonally into one
// 128-row tile per (window pair, head).
//
//   rows   0..48  = window A,  rows 64..112 = window B  (rows 49..63 / 113..127 are zero padding)
//   S[128x128] = Q K^T            one tcgen05.mma pair (K = 32), fp32 in TMEM; only the two 49x49 diagonal blocks
//                                 are read back
//   softmax                       one thread per row: tcgen05.ld of its window's 64 columns, scale + bias + shift
//                                 mask, exp, sum in registers; P written 16-bit into shared memory in the 128-byte
//                                 swizzled K-major layout, off-diagonal blocks stay zero
//   O[128x32]  = P V              eight tcgen05.mma (K = 16 tokens each) with V used in place as an MN-major operand
//                                 (TMA delivers [token][dim] rows with the 64-byte swizzle = the canonical layout)
//
// Q, K, V arrive by TMA straight from the window-ordered qkv tensor (49-row x 64-byte boxes, two per operand),
// double-buffered across work items; the relative-position bias of the CTA's head lives in shared memory.
// Replaces the same reference ops as attention.cu (HF:swin/modeling_swin.py:424-455, 556-582) with the same
// operation order: S/sqrt(32) + bias + mask(-100) -> softmax -> P V.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = softmax / epilogue (one TMEM lane
// quadrant each).  160 TMEM columns and ~90 KB of shared memory per CTA, so two CTAs share an SM and one CTA's
// MMAs overlap the other's softmax.
#include <cudaTypedefs.h>

#include <type_traits>

#include "errors.h"
#include "gemm.cuh"
#include "rowops.cuh"

namespace csvit {

constexpr int TA_L = 49;
constexpr int TA_THREADS = 192;
constexpr uint32_t TA_OPER_BYTES = 128 * 64;            // one operand tile: 128 rows x 64 B
constexpr uint32_t TA_STAGE_BYTES = 3 * TA_OPER_BYTES;  // Q, K, V
constexpr uint32_t TA_P_BYTES = 2 * 128 * 128;          // P: two 64-token slabs of 128 rows x 128 B
constexpr uint32_t TA_BIAS_BYTES = ((TA_L * TA_L * 4 + 127) / 128) * 128;
constexpr size_t TA_SMEM = 1024 + 2 * TA_STAGE_BYTES + TA_P_BYTES + TA_BIAS_BYTES + 256 + 256;

// K-major operand with 64-byte rows, 64-byte swizzle: 8-row atoms of 512 B.
__device__ __forceinline__ uint64_t make_sw64_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(512 >> 4) << 32;     // SBO: next 8 rows
  d |= static_cast<uint64_t>(1) << 46;            // sm_100 descriptor version
  d |= static_cast<uint64_t>(4) << 61;            // SWIZZLE_64B
  return d;
}
// MN-major operand: [K rows][32 MN elements = 64 B] with the 64-byte swizzle; 8 K-rows per 512-byte atom.
__device__ __forceinline__ uint64_t make_sw64_mnmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(512 >> 4) << 16;     // LBO: stride between MN repeats (only one 64-byte span here)
  d |= static_cast<uint64_t>(512 >> 4) << 32;     // SBO: next group of 8 K-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_ex(uint32_t fmt, int M, int N, uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (b_mn_major << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

template <int FMT>   // 0 = fp16, 1 = bf16
__global__ void __launch_bounds__(TA_THREADS, 2)
win_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const float* __restrict__ bias_plain, void* __restrict__ out_,
                   int num_windows, int C, int heads, WinGeom g, int nW, float scale) {
  using T16 = typename std::conditional<FMT == 1, __nv_bfloat16, __half>::type;
  T16* out = static_cast<T16*>(out_);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t p_off = 2 * TA_STAGE_BYTES, bias_off = p_off + TA_P_BYTES, reg_off = bias_off + TA_BIAS_BYTES;
  float* bias_s = reinterpret_cast<float*>(smem + bias_off);
  int8_t* region_all = reinterpret_cast<int8_t*>(smem + reg_off);        // [2 items in flight][2 windows][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + reg_off + 256);
  uint64_t* in_full = bars;        // [2]
  uint64_t* in_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % heads;
  const int num_pairs = (num_windows + 1) >> 1;
  const int pstart = blockIdx.x / heads, pstride = gridDim.x / heads;

  // zero the operand stages and P once: padding rows / off-diagonal blocks are never written afterwards
  for (uint32_t i = threadIdx.x; i < (2 * TA_STAGE_BYTES + TA_P_BYTES) / 16; i += TA_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < TA_L * TA_L; i += TA_THREADS) bias_s[i] = __ldg(bias_plain + h * TA_L * TA_L + i);
  for (int i = threadIdx.x; i < 256; i += TA_THREADS) region_all[i] = 0;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    for (int s = 0; s < 2; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_O = tmem_base + 128;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int p = pstart; p < num_pairs; p += pstride, ++it) {
        const int st = it & 1;
        mbar_wait(&in_empty[st], ((it >> 1) & 1u) ^ 1u);
        const int nwin = (2 * p + 1 < num_windows) ? 2 : 1;
        mbar_arrive_expect_tx(&in_full[st], uint32_t(nwin) * 3u * TA_L * 64u);
        uint8_t* sb = smem + size_t(st) * TA_STAGE_BYTES;
        for (int wdx = 0; wdx < nwin; ++wdx) {
          const int row0 = (2 * p + wdx) * TA_L;
          for (int op = 0; op < 3; ++op)   // Q, K, V column blocks of the qkv tensor
            tma_load_2d(sb + op * TA_OPER_BYTES + wdx * 64 * 64, &tmQKV, &in_full[st], op * C + h * 32, row0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_ex(uint32_t(FMT), 128, 128, 0);
      constexpr uint32_t idesc_o = make_idesc_ex(uint32_t(FMT), 128, 32, 1);
      uint32_t it = 0;
      for (int p = pstart; p < num_pairs; p += pstride, ++it) {
        const int st = it & 1;
        const uint32_t sb = base + uint32_t(st) * TA_STAGE_BYTES;
        mbar_wait(&in_full[st], (it >> 1) & 1u);
        tc_fence_after();
        // S = Q K^T : 2 k-steps of 16 (32 bytes each inside the 64-byte rows).  The previous item's softmax has
        // finished reading S (its p_full was awaited before its PV was issued).
        const uint64_t qd = make_sw64_kmajor_desc(sb), kd = make_sw64_kmajor_desc(sb + TA_OPER_BYTES);
        umma_ss<false>(tm_S, qd, kd, idesc_s, 0u);
        umma_ss<false>(tm_S, qd + 2, kd + 2, idesc_s, 1u);
        umma_commit(s_full);
        // O = P V once the softmax warps have published P (and drained the previous O)
        mbar_wait(p_full, it & 1u);
        mbar_wait(o_empty, (it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t vb = sb + 2 * TA_OPER_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t pd = make_sw128_kmajor_desc(base + p_off + uint32_t(ks >> 2) * (128 * 128) + uint32_t(ks & 3) * 32);
          const uint64_t vd = make_sw64_mnmajor_desc(vb + uint32_t(ks) * 1024);
          umma_ss<false>(tm_O, pd, vd, idesc_o, ks ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&in_empty[st]);
      }
    }
  } else {
    // ---------------- softmax + epilogue: one thread per tile row ----------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // tile row
    const int wd = r >> 6;                   // window of the pair
    const int i = r & 63;                    // slot inside the window (valid if < 49)
    const bool bf = FMT == 1;
    const int nWy = g.H / g.ws;
    uint32_t it = 0;
    int prev_p = -1;
    auto store_prev = [&](uint32_t pit) {
      // O of the previous item: TMEM -> 16 bit -> 64 contiguous bytes of the head's column block
      mbar_wait(o_full, pit & 1u);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32(tm_O + (uint32_t(quad * 32) << 16), o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int wg = 2 * prev_p + wd;
      if (i < TA_L && wg < num_windows) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(wg) * TA_L + i) * C + h * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(pack16(bf, __uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])),
                              pack16(bf, __uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])),
                              pack16(bf, __uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])),
                              pack16(bf, __uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])));
      }
    };
    for (int p = pstart; p < num_pairs; p += pstride, ++it) {
      const int wg = 2 * p + wd;
      const int w = wg % nW;
      const int wy = w / g.nWx, wx = w - wy * g.nWx;
      const bool masked = g.shift > 0 && wg < num_windows && (wy == nWy - 1 || wx == g.nWx - 1);
      // region ids of this pair's windows (each window is served by two warps; double-buffered across items because
      // a warp may run one item ahead of its neighbour)
      int8_t* region_s = region_all + (it & 1u) * 128;
      if (i < TA_L) region_s[wd * 64 + i] = masked ? static_cast<int8_t>(win_region(g, w, i)) : int8_t(0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, it & 1u);
      tc_fence_after();
      uint32_t pk[32];
      {
        uint32_t sv[64];
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
        tmem_ld_32x32(tm_S + (uint32_t(quad * 32) << 16) + uint32_t(wd * 64), lo);
        tmem_ld_32x32(tm_S + (uint32_t(quad * 32) << 16) + uint32_t(wd * 64 + 32), hi);
        tmem_ld_wait();
        const int ib = i < TA_L ? i : TA_L - 1;
        const float* brow = bias_s + ib * TA_L;
        const int8_t* reg = region_s + wd * 64;
        const int myreg = reg[ib];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float v = -INFINITY;
          if (j < TA_L) {
            v = fmaf(__uint_as_float(sv[j]), scale, brow[j]);
            if (masked && reg[j] != myreg) v += -100.0f;
          }
          sv[j] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float e = j < TA_L ? __expf(__uint_as_float(sv[j]) - mx) : 0.f;
          sv[j] = __float_as_uint(e);
          sum += e;
        }
        const float inv = (i < TA_L) ? 1.0f / sum : 0.f;   // padding rows publish an all-zero P row
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = pack16(bf, __uint_as_float(sv[2 * j]) * inv, __uint_as_float(sv[2 * j + 1]) * inv);
      }
      if (prev_p >= 0) store_prev(it - 1);      // also guarantees PV(it-1) has finished reading P
      uint8_t* prow = smem + p_off + wd * (128 * 128) + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      prev_p = p;
    }
    if (prev_p >= 0) store_prev(it - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int FMT>
static int launch_ta(const CUtensorMap& tm, const float* bias_plain, void* out, int num_windows, int C, int heads,
                     const WinGeom& g, int nW, cudaStream_t stream) {
  static bool configured = false;
  auto kern = win_attn_tc_kernel<FMT>;
  if (!configured) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TA_SMEM)));
    configured = true;
  }
  const int num_pairs = (num_windows + 1) / 2;
  int per_head = num_pairs;
  const int cap = (2 * num_sms()) / heads > 0 ? (2 * num_sms()) / heads : 1;
  if (per_head > cap) per_head = cap;
  kern<<<per_head * heads, TA_THREADS, TA_SMEM, stream>>>(tm, bias_plain, out, num_windows, C, heads, g, nW, 0.17677669529663687f);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_window_attention_tc(const void* qkv, const float* bias_plain, void* out, int dtype, int B, int H, int W, int C,
                               int heads, int ws, int shift, cudaStream_t stream) {
  CSVIT_REQUIRE(ws == 7, "window_attention(tcgen05): only window 7 is built (got %d)", ws);
  CSVIT_REQUIRE(C == heads * 32, "window_attention(tcgen05): head_dim must be 32 (C=%d heads=%d)", C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0, "window_attention: %dx%d not divisible by window %d", H, W, ws);
  const int nW = (H / ws) * (W / ws);
  const long long windows = static_cast<long long>(B) * nW;
  if (windows <= 0) return 0;
  CSVIT_REQUIRE(windows < (1ll << 30), "window_attention: too many windows");
  // 49-row x 32-column (64-byte) boxes with the 64-byte swizzle
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
  }
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (3 * C * 2) % 16 == 0, "window_attention: qkv must be 16-byte aligned");
  CUtensorMap tm;
  cuuint64_t gdim[2] = {cuuint64_t(3 * C), cuuint64_t(windows * TA_L)};
  cuuint64_t gstr[1] = {cuuint64_t(3 * C) * 2};
  cuuint32_t box[2] = {32, TA_L};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(qkv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(qkv) failed with CUresult %d", int(r));
  WinGeom g = make_geom(H, W, ws, shift);
  if (dtype == DT_BF16) return launch_ta<1>(tm, bias_plain, out, static_cast<int>(windows), C, heads, g, nW, stream);
  return launch_ta<0>(tm, bias_plain, out, static_cast<int>(windows), C, heads, g, nW, stream);
}

}  // namespace csvit
