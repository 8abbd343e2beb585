// Swin (shifted-)window attention core on 16-bit Q/K/V with tensor-core MMAs and fp32 softmax.
//
// Replaces K5-K9 of SURVEY.md §2.3: the batched 49x32x49 `bmm`, the /sqrt(d), the relative-position-bias
// gather+add, the shift-mask build+add, `softmax`, the second `bmm` and the head-merge permute copy
// (HF:swin/modeling_swin.py:424-455, 556-582).  Order of operations kept:
//   S = QK^T / sqrt(32) -> + bias[h,i,j] -> + mask(0 / -100, additive, not -inf) -> softmax_j -> P V.
// The mask is never materialised: region ids come from the closed form in common.cuh, and only windows in
// the last window row / column of a shifted block evaluate it at all.
//
// Work decomposition: ONE WARP per (window, head), one head per CTA.  The warp pulls its 49x32 Q, K and V slices with cp.async
// into a private shared-memory region, then walks the four 16-row query tiles entirely in registers
// (mma.sync m16n8k16, P re-used as the A operand of P.V).  No block-level barrier exists, so a CTA's four
// warps and the three CTAs per SM are always at different points of load / compute and the HBM stream
// (the kernel's bound: 8C bytes per token) stays busy.  The bias table is pre-swizzled into the accumulator
// fragment layout (csvit_expand_rel_bias_mma) so each lane fetches it with one coalesced 16-byte load per
// 16x8 tile, with -inf already in the padding columns.
#include <type_traits>

#include "errors.h"
#include "rowops.cuh"

namespace csvit {

// Per-warp shared memory: Q, K, V as 49 dense 64-byte rows each (16-byte chunks XOR-swizzled by (row>>1)&3 so that
// ldmatrix is conflict-free without padding), then 15 zero rows.  Fragment loads that run past a 49-row matrix
// alias into the next one: harmless for Q (those query rows are never stored) and K (those key columns carry a
// -inf bias); for V they hit the zero rows, and P is exactly 0 there.
constexpr int WA_L = 49;
constexpr int WA_ROW = 64;                                    // bytes per row (32 x 16 bit)
constexpr int WA_WARP_BYTES = (3 * WA_L + 15) * WA_ROW + 64 + 128;  // + region ids + token ids of the window's 49 slots
constexpr int WA_WARPS = 4;
constexpr int WA_BIAS_BYTES = 4 * 7 * 32 * 16;                // one head's fragment-ordered bias table
constexpr float WA_LOG2E = 1.4426950408889634f;
// Softmax in the log2 domain: log2(e) is folded into the logit scale, the bias table (csvit_expand_rel_bias_mma) and the mask
// value, so each probability costs one FADD and one MUFU.EX2 (expf's range fix-up alone is four more instructions).
__device__ __forceinline__ float wa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> struct WaOnes;
template <> struct WaOnes<__nv_bfloat16> { static constexpr uint32_t v = 0x3F803F80u; };
template <> struct WaOnes<__half> { static constexpr uint32_t v = 0x3C003C00u; };
__device__ __forceinline__ uint32_t wa_off(int row, int chunk) { return uint32_t(row * WA_ROW + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// bias_frag[h][mt][j][lane] = float4 of the accumulator fragment (c0..c3) of query tile mt / key tile j:
//   c0,c1 -> row 16mt + lane/4,     cols 8j + 2(lane%4) + {0,1};   c2,c3 -> row + 8, same cols.
// Padding columns (>= 49) hold -inf so they vanish in the softmax, padding rows hold 0.
__global__ void expand_rel_bias_mma_kernel(const float* __restrict__ table, float* __restrict__ out, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = heads * 4 * 7 * 32 * 4;
  if (idx >= total) return;
  const int e = idx & 3, lane = (idx >> 2) & 31;
  int t = idx >> 7;
  const int j = t % 7; t /= 7;
  const int mt = t & 3, h = t >> 2;
  const int row = 16 * mt + (lane >> 2) + ((e >> 1) ? 8 : 0);
  const int col = 8 * j + 2 * (lane & 3) + (e & 1);
  float v;
  if (col >= WA_L) v = -INFINITY;
  else if (row >= WA_L) v = 0.f;
  else v = table[rel_pos_index(7, row, col) * heads + h] * WA_LOG2E;   // the kernel's logits live in the log2 domain
  out[idx] = v;
}

// qkv: window-ordered tokens [B*N, 3C] (Q | K | V column blocks), out: [B*N, C] window-ordered.
// A CTA serves ONE head (gridDim.x is a multiple of `heads`): its bias table sits in shared memory, its four warps
// take windows round-robin.
template <typename T>
__global__ void __launch_bounds__(WA_WARPS * 32, 4)
win_attn_warp_kernel(const T* __restrict__ qkv, const float4* __restrict__ bias_frag, T* __restrict__ out, int num_windows,
                     int C, int heads, WinGeom g, int nW, float scale, int tok_order) {
  extern __shared__ __align__(16) uint8_t wa_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* bias_s = reinterpret_cast<float4*>(wa_smem);
  uint8_t* Qs = wa_smem + WA_BIAS_BYTES + warp * WA_WARP_BYTES;
  uint8_t* Ks = Qs + WA_L * WA_ROW;
  uint8_t* Vs = Ks + WA_L * WA_ROW;
  int8_t* region_s = reinterpret_cast<int8_t*>(Vs + (WA_L + 15) * WA_ROW);
  int16_t* tok_s = reinterpret_cast<int16_t*>(region_s + 64);   // tok_order: token id (within the image) of each window slot

  const int h = blockIdx.x % heads;
  for (int i = threadIdx.x; i < WA_BIAS_BYTES / 16; i += WA_WARPS * 32) bias_s[i] = __ldg(bias_frag + h * (WA_BIAS_BYTES / 16) + i);
  for (int i = lane; i < (WA_WARP_BYTES >> 2); i += 32) reinterpret_cast<uint32_t*>(Qs)[i] = 0u;
  __syncthreads();

  const int ld_qkv = 3 * C;
  const int nWy = g.H / g.ws;
  const int wstride = (gridDim.x / heads) * WA_WARPS;
  for (int wg = (blockIdx.x / heads) * WA_WARPS + warp; wg < num_windows; wg += wstride) {   // global window = b*nW + w
    const int w = wg % nW;
    const long long row0 = static_cast<long long>(wg) * WA_L;
    const T* src = qkv + row0 * ld_qkv + h * 32;
#pragma unroll
    for (int t = 0; t < 19; ++t) {
      const int idx = lane + 32 * t;
      if (idx < 3 * WA_L * 4) {
        const int which = idx / (WA_L * 4), rem = idx - which * (WA_L * 4);
        const int r = rem >> 2, ch = rem & 3;
        cp_async16(Qs + which * (WA_L * WA_ROW) + wa_off(r, ch), src + static_cast<long long>(r) * ld_qkv + which * C + ch * 8);
      }
    }
    const int wy = w / g.nWx, wx = w - wy * g.nWx;
    const bool masked = g.shift > 0 && (wy == nWy - 1 || wx == g.nWx - 1);   // warp-uniform
    if (masked) {
      region_s[lane] = static_cast<int8_t>(win_region(g, w, lane));
      if (lane + 32 < WA_L) region_s[lane + 32] = static_cast<int8_t>(win_region(g, w, lane + 32));
    }
    long long orow0 = row0;          // window-ordered output: row0 + slot;  token-ordered: image base + token id of the slot
    if (tok_order) {                 // window_reverse + roll(+shift) folded into the store, so the out-proj GEMM sees plain rows
      tok_s[lane] = static_cast<int16_t>(win_row_to_token(g, w * WA_L + lane));
      if (lane + 32 < WA_L) tok_s[lane + 32] = static_cast<int16_t>(win_row_to_token(g, w * WA_L + lane + 32));
      orow0 = static_cast<long long>(wg / nW) * g.N;
    }
    cp_async_wait_all();
    __syncwarp();

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int m0 = mt * 16;
      // ---- S = Q K^T ----
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldsm_x4(qa[ks], Qs + wa_off(m0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
      float s[7][4];
      const float4* bf = bias_s + (mt * 7) * 32 + lane;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        uint32_t kb[4];
        ldsm_x4(kb, Ks + wa_off(j * 8 + (lane & 7), lane >> 3));
        mma_16816<T>(s[j], qa[0], kb[0], kb[1]);
        mma_16816<T>(s[j], qa[1], kb[2], kb[3]);
      }
      // ---- scale + bias (+ mask), softmax over the 49 keys (fp32) ----
      const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float4 b = bf[j * 32];
        s[j][0] = fmaf(s[j][0], scale, b.x);
        s[j][1] = fmaf(s[j][1], scale, b.y);
        s[j][2] = fmaf(s[j][2], scale, b.z);
        s[j][3] = fmaf(s[j][3], scale, b.w);
      }
      if (masked) {
        const int reg0 = region_s[r0 < WA_L ? r0 : WA_L - 1], reg1 = region_s[r1 < WA_L ? r1 : WA_L - 1];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = j * 8 + (lane & 3) * 2 + e;
            const int rc = region_s[c < WA_L ? c : WA_L - 1];
            if (rc != reg0) s[j][e] += -100.0f * WA_LOG2E;
            if (rc != reg1) s[j][2 + e] += -100.0f * WA_LOG2E;
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      // ---- O = P V with unnormalised P re-used in registers as the A operand; a ones column (o[4]) yields the row sums of the
      //      ROUNDED probabilities on the tensor core, so the weights that multiply V sum to exactly one ----
      float o[5][4];
#pragma unroll
      for (int n = 0; n < 5; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = Half16<T>::pack(wa_ex2(s[2 * kk][0] - mx0), wa_ex2(s[2 * kk][1] - mx0));
        pa[1] = Half16<T>::pack(wa_ex2(s[2 * kk][2] - mx1), wa_ex2(s[2 * kk][3] - mx1));
        if (2 * kk + 1 < 7) {
          pa[2] = Half16<T>::pack(wa_ex2(s[2 * kk + 1][0] - mx0), wa_ex2(s[2 * kk + 1][1] - mx0));
          pa[3] = Half16<T>::pack(wa_ex2(s[2 * kk + 1][2] - mx1), wa_ex2(s[2 * kk + 1][3] - mx1));
        } else {
          pa[2] = 0u; pa[3] = 0u;
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t vb[4];
          ldsm_x4_t(vb, Vs + wa_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)));
          mma_16816<T>(o[2 * np], pa, vb[0], vb[1]);
          mma_16816<T>(o[2 * np + 1], pa, vb[2], vb[3]);
        }
        mma_16816<T>(o[4], pa, WaOnes<T>::v, WaOnes<T>::v);
      }
      const float inv0 = 1.0f / o[4][0], inv1 = 1.0f / o[4][2];
      // ---- store (head merge folded into the column offset) ----
      int or0 = r0, or1 = r1;
      if (tok_order) { or0 = tok_s[r0 < WA_L ? r0 : 0]; or1 = tok_s[r1 < WA_L ? r1 : 0]; }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int col = h * 32 + n * 8 + (lane & 3) * 2;
        if (r0 < WA_L) *reinterpret_cast<uint32_t*>(out + (orow0 + or0) * C + col) = Half16<T>::pack(o[n][0] * inv0, o[n][1] * inv0);
        if (r1 < WA_L) *reinterpret_cast<uint32_t*>(out + (orow0 + or1) * C + col) = Half16<T>::pack(o[n][2] * inv1, o[n][3] * inv1);
      }
    }
    __syncwarp();  // all lanes are done with this window's tiles before the next cp.async overwrites them
  }
}

int launch_expand_rel_bias_mma(const float* table, float* out, int heads, cudaStream_t stream) {
  const int total = heads * 4 * 7 * 32 * 4;
  expand_rel_bias_mma_kernel<<<(total + 255) / 256, 256, 0, stream>>>(table, out, heads);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int launch_wa(const void* qkv, const float* bias_frag, void* out, int num_windows, int C, int heads, const WinGeom& g,
                     int nW, int tok_order, cudaStream_t stream) {
  static bool configured = false;
  auto kern = win_attn_warp_kernel<T>;
  const int smem = WA_BIAS_BYTES + WA_WARPS * WA_WARP_BYTES;
  if (!configured) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  // one head per CTA: grid is a multiple of `heads`, about 4 CTAs per SM
  int per_head = (num_windows + WA_WARPS - 1) / WA_WARPS;
  const int cap = (148 * 4) / heads > 0 ? (148 * 4) / heads : 1;
  if (per_head > cap) per_head = cap;
  kern<<<per_head * heads, WA_WARPS * 32, smem, stream>>>(static_cast<const T*>(qkv), reinterpret_cast<const float4*>(bias_frag),
                                                          static_cast<T*>(out), num_windows, C, heads, g, nW,
                                                          0.17677669529663687f * WA_LOG2E, tok_order);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_window_attention_mma(const void* qkv, const float* bias_frag, void* out, int dtype, int B, int H, int W,
                                int C, int heads, int ws, int shift, int tok_order, cudaStream_t stream) {
  CSVIT_REQUIRE(ws == 7, "window_attention(16-bit): only window 7 is built (got %d)", ws);
  CSVIT_REQUIRE(C == heads * 32, "window_attention(16-bit): head_dim must be 32 (C=%d heads=%d)", C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0, "window_attention: %dx%d not divisible by window %d", H, W, ws);
  CSVIT_REQUIRE((reinterpret_cast<uintptr_t>(bias_frag) & 15) == 0, "window_attention: bias table must be 16-byte aligned");
  const int nW = (H / ws) * (W / ws);
  const long long items = static_cast<long long>(B) * nW * heads;
  if (items <= 0) return 0;
  CSVIT_REQUIRE(items < (1ll << 31), "window_attention: too many work items");
  WinGeom g = make_geom(H, W, ws, shift);
  CSVIT_REQUIRE(H * W < 32768, "window_attention: %dx%d tokens per image exceed the 16-bit token table", H, W);
  if (dtype == DT_BF16) return launch_wa<__nv_bfloat16>(qkv, bias_frag, out, B * nW, C, heads, g, nW, tok_order, stream);
  return launch_wa<__half>(qkv, bias_frag, out, B * nW, C, heads, g, nW, tok_order, stream);
}

}  // namespace csvit
