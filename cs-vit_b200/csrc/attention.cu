// Attention cores.
//
// (1) win_attn_mma_kernel - Swin (shifted-)window attention on bf16 Q/K/V with tensor-core MMAs, fp32
//     softmax.  Replaces K5-K9 of SURVEY.md §2.3: the batched 49x32x49 `bmm`, the /sqrt(d), the
//     relative-position-bias gather+add, the shift-mask build+add, `softmax`, the second `bmm` and the
//     head-merge permute copy (HF:swin/modeling_swin.py:424-455, 556-582).  Order of operations kept:
//     S = QK^T / sqrt(32) -> + bias[h,i,j] -> + mask(0 / -100, additive, not -inf) -> softmax_j -> P V.
//     The mask is never materialised: region ids come from the closed form in common.cuh.
// (2) attention_simt_kernel - exact-fp32 dense attention for short sequences (<= 64 keys, head_dim 32):
//     the CS-ViT head's MHA (ref:cs_vit/net/transformer_module.py:250-282) whose logits are MULTIPLIED by
//     sqrt(head_dim) (line 273, quirk Q1: near-argmax softmax, kept in fp32 on purpose), and the fp32
//     validation mode of (1).
#include <type_traits>

#include "errors.h"
#include "rowops.cuh"

namespace csvit {

// ----------------------------------------------------------------------------------------------------
// (1) window attention, mma.sync m16n8k16 bf16
// ----------------------------------------------------------------------------------------------------
constexpr int WA_LD = 40;  // smem row pitch in bf16 (80 B): conflict-free ldmatrix

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
  if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// One CTA (4 warps) per (window, head) work item, grid-stride.  Warp w owns query rows 16w..16w+15.
// qkv: window-ordered tokens [B*N, 3C] (Q | K | V column blocks), out: [B*N, C] window-ordered.
template <typename T>
__global__ void __launch_bounds__(128)
win_attn_mma_kernel(const T* __restrict__ qkv, const float* __restrict__ bias_exp,
                    T* __restrict__ out, int num_items, int C, int heads, WinGeom g, int nW, float scale) {
  constexpr int L = 49;
  __shared__ __align__(16) T Qs[64 * WA_LD];
  __shared__ __align__(16) T Ks[64 * WA_LD];
  __shared__ __align__(16) T Vs[64 * WA_LD];
  __shared__ int region_s[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * WA_LD / 2; i += 128) {
    reinterpret_cast<uint32_t*>(Qs)[i] = 0u;
    reinterpret_cast<uint32_t*>(Ks)[i] = 0u;
    reinterpret_cast<uint32_t*>(Vs)[i] = 0u;
  }
  if (tid < 64) region_s[tid] = 0;

  const int ld_qkv = 3 * C;
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int h = item % heads;
    const int wg = item / heads;      // global window index = b*nW + w
    const int w = wg % nW;
    const long long row0 = static_cast<long long>(wg) * L;
    __syncthreads();                  // previous item fully consumed (and the zero fill on the first pass)
    for (int idx = tid; idx < 3 * L * 4; idx += 128) {
      int which = idx / (L * 4), rem = idx - which * (L * 4);
      int r = rem >> 2, ch = rem & 3;
      const uint4 v = *reinterpret_cast<const uint4*>(qkv + (row0 + r) * ld_qkv + which * C + h * 32 + ch * 8);
      T* dst = which == 0 ? Qs : (which == 1 ? Ks : Vs);
      *reinterpret_cast<uint4*>(dst + r * WA_LD + ch * 8) = v;
    }
    if (g.shift > 0 && tid < L) region_s[tid] = win_region(g, w, tid);
    __syncthreads();

    const int m0 = warp * 16;
    // ---- S = Q K^T ----
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldsm_x4(qa[ks], Qs + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * WA_LD + ks * 16 + (lane >> 4) * 8);
    float s[7][4];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t kb[4];
      ldsm_x4(kb, Ks + (j * 8 + (lane & 7)) * WA_LD + (lane >> 3) * 8);
      mma_16816<T>(s[j], qa[0], kb[0], kb[1]);
      mma_16816<T>(s[j], qa[1], kb[2], kb[3]);
    }
    // ---- scale + bias + mask, softmax over the 49 keys (fp32) ----
    const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
    const int rb0 = r0 < L ? r0 : L - 1, rb1 = r1 < L ? r1 : L - 1;
    const float* b0p = bias_exp + (static_cast<long long>(h) * L + rb0) * L;
    const float* b1p = bias_exp + (static_cast<long long>(h) * L + rb1) * L;
    const int reg0 = region_s[rb0], reg1 = region_s[rb1];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = j * 8 + (lane & 3) * 2 + e;
        if (c < L) {
          const int rc = region_s[c];
          s[j][e] = s[j][e] * scale + __ldg(b0p + c) + (rc != reg0 ? -100.0f : 0.0f);
          s[j][2 + e] = s[j][2 + e] * scale + __ldg(b1p + c) + (rc != reg1 ? -100.0f : 0.0f);
        } else {
          s[j][e] = -INFINITY;
          s[j][2 + e] = -INFINITY;
        }
        mx0 = fmaxf(mx0, s[j][e]);
        mx1 = fmaxf(mx1, s[j][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[j][e] = __expf(s[j][e] - mx0);
        s[j][2 + e] = __expf(s[j][2 + e] - mx1);
        sum0 += s[j][e];
        sum1 += s[j][2 + e];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    // ---- O = P V  (P re-used in registers as the A operand) ----
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = Half16<T>::pack(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
      pa[1] = Half16<T>::pack(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
      if (2 * kk + 1 < 7) {
        pa[2] = Half16<T>::pack(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
        pa[3] = Half16<T>::pack(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
      } else {
        pa[2] = 0u; pa[3] = 0u;
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vb[4];
        ldsm_x4_t(vb, Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * WA_LD + np * 16 + (lane >> 4) * 8);
        mma_16816<T>(o[2 * np], pa, vb[0], vb[1]);
        mma_16816<T>(o[2 * np + 1], pa, vb[2], vb[3]);
      }
    }
    // ---- store (head merge folded into the column offset) ----
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = h * 32 + n * 8 + (lane & 3) * 2;
      if (r0 < L) *reinterpret_cast<uint32_t*>(out + (row0 + r0) * C + col) = Half16<T>::pack(o[n][0], o[n][1]);
      if (r1 < L) *reinterpret_cast<uint32_t*>(out + (row0 + r1) * C + col) = Half16<T>::pack(o[n][2], o[n][3]);
    }
  }
}

int launch_window_attention_mma(const void* qkv, const float* bias_exp, void* out, int dtype, int B, int H, int W,
                                int C, int heads, int ws, int shift, cudaStream_t stream) {
  CSVIT_REQUIRE(ws == 7, "window_attention(bf16): only window 7 is built (got %d)", ws);
  CSVIT_REQUIRE(C == heads * 32, "window_attention(bf16): head_dim must be 32 (C=%d heads=%d)", C, heads);
  CSVIT_REQUIRE(H % ws == 0 && W % ws == 0, "window_attention: %dx%d not divisible by window %d", H, W, ws);
  const int nW = (H / ws) * (W / ws);
  const long long items = static_cast<long long>(B) * nW * heads;
  if (items <= 0) return 0;
  CSVIT_REQUIRE(items < (1ll << 31), "window_attention: too many work items");
  WinGeom g = make_geom(H, W, ws, shift);
  int blocks = static_cast<int>(items < 148 * 12 ? items : 148 * 12);
  const float scale = 0.17677669529663687f;  // 1/sqrt(32)
  if (dtype == DT_BF16)
    win_attn_mma_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(qkv), bias_exp,
        static_cast<__nv_bfloat16*>(out), static_cast<int>(items), C, heads, g, nW, scale);
  else
    win_attn_mma_kernel<__half><<<blocks, 128, 0, stream>>>(static_cast<const __half*>(qkv), bias_exp,
        static_cast<__half*>(out), static_cast<int>(items), C, heads, g, nW, scale);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------------------------------
// (2) exact fp32 dense attention for short sequences
// ----------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

constexpr int SA_MAXS = 64;
constexpr int SA_HD = 32;

// One CTA (4 warps) per (sequence, head); warp per query row, lane per key (two keys per lane).
template <typename T>
__global__ void __launch_bounds__(128)
attention_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ out,
                      long long ldq, long long ldk, long long ldv, long long ldo, int num_items, int Lq, int S, int heads,
                      float scale, const float* __restrict__ bias, WinGeom g, int nW) {
  __shared__ float Ks[SA_MAXS][SA_HD + 1];
  __shared__ float Vs[SA_MAXS][SA_HD + 1];
  __shared__ float Qs[4][SA_HD];
  __shared__ float Ps[4][SA_MAXS];
  __shared__ int region_s[SA_MAXS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int h = item % heads, seq = item / heads;
    __syncthreads();
    for (int idx = tid; idx < S * SA_HD; idx += 128) {
      int r = idx >> 5, d = idx & 31;
      Ks[r][d] = to_f<T>(k[(static_cast<long long>(seq) * S + r) * ldk + h * SA_HD + d]);
      Vs[r][d] = to_f<T>(v[(static_cast<long long>(seq) * S + r) * ldv + h * SA_HD + d]);
    }
    if (tid < SA_MAXS) region_s[tid] = (g.shift > 0 && tid < S) ? win_region(g, seq % nW, tid) : 0;
    __syncthreads();
    for (int i = warp; i < Lq; i += 4) {
      Qs[warp][lane] = to_f<T>(q[(static_cast<long long>(seq) * Lq + i) * ldq + h * SA_HD + lane]);
      __syncwarp();
      float sc[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        float a = -INFINITY;
        if (j < S) {
          a = 0.f;
#pragma unroll
          for (int d = 0; d < SA_HD; ++d) a = fmaf(Qs[warp][d], Ks[j][d], a);
          a *= scale;
          if (bias) a += __ldg(bias + (static_cast<long long>(h) * Lq + i) * S + j);
          if (g.shift > 0 && region_s[j] != region_s[i]) a += -100.0f;
        }
        sc[half] = a;
      }
      const float mx = warp_max(fmaxf(sc[0], sc[1]));
      const float e0 = lane < S ? expf(sc[0] - mx) : 0.f;
      const float e1 = lane + 32 < S ? expf(sc[1] - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      Ps[warp][lane] = e0 * inv;
      Ps[warp][lane + 32] = e1 * inv;
      __syncwarp();
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(Ps[warp][j], Vs[j][lane], acc);
      out[(static_cast<long long>(seq) * Lq + i) * ldo + h * SA_HD + lane] = from_f<T>(acc);
      __syncwarp();
    }
  }
}

int launch_attention_simt(const void* q, const void* k, const void* v, void* out, int dtype, long long ldq, long long ldk,
                          long long ldv, long long ldo, int n_seq, int Lq, int S, int heads, float scale,
                          const float* bias, int mH, int mW, int mws, int mshift, cudaStream_t stream) {
  CSVIT_REQUIRE(S >= 1 && S <= SA_MAXS, "attention_simt: key length %d outside [1,%d]", S, SA_MAXS);
  CSVIT_REQUIRE(Lq >= 1, "attention_simt: empty query");
  const long long items = static_cast<long long>(n_seq) * heads;
  if (items <= 0) return 0;
  CSVIT_REQUIRE(items < (1ll << 31), "attention_simt: too many work items");
  WinGeom g = make_geom(mH > 0 ? mH : 1, mW > 0 ? mW : 1, mws > 0 ? mws : 1, mshift);
  int nW = mshift > 0 ? (mH / mws) * (mW / mws) : 1;
  if (mshift > 0) CSVIT_REQUIRE(S == mws * mws && Lq == S, "attention_simt: window mask needs Lq == S == ws^2");
  int blocks = static_cast<int>(items < 148 * 16 ? items : 148 * 16);
  if (dtype == DT_F16)
    attention_simt_kernel<__half><<<blocks, 128, 0, stream>>>(
        static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v),
        static_cast<__half*>(out), ldq, ldk, ldv, ldo, static_cast<int>(items), Lq, S, heads, scale, bias, g, nW);
  else if (dtype == DT_BF16)
    attention_simt_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v),
        static_cast<__nv_bfloat16*>(out), ldq, ldk, ldv, ldo, static_cast<int>(items), Lq, S, heads, scale, bias, g, nW);
  else
    attention_simt_kernel<float><<<blocks, 128, 0, stream>>>(
        static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v), static_cast<float*>(out),
        ldq, ldk, ldv, ldo, static_cast<int>(items), Lq, S, heads, scale, bias, g, nW);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace csvit
