// GEMM engine interface shared by the tcgen05 kernel, the SIMT fp32 kernel and the C ABI.
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )        (both operands K-major, i.e. nn.Linear layout)
#pragma once
#include "common.cuh"

namespace csvit {

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum : int { ROWMAP_IDENTITY = 0, ROWMAP_WINDOW = 1 };
enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
__host__ __device__ inline int dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// What happens to one accumulator row-chunk after the MMA:
//   v = acc (+ bias[col]) -> act -> (+ resid[orow, col]) -> out[orow, col]  (fp32, bf16 or fp16)
// orow = row for ROWMAP_IDENTITY; for ROWMAP_WINDOW the GEMM rows are window-ordered tokens and orow is the
// original token row (window_reverse + roll(+shift) folded into the store address).
struct EpiParams {
  const float* bias;
  const float* resid;
  void* out;
  long long ldo, ldr;
  int out_dtype;
  int act;
  int M, N;
  int vec_ok;     // 1: rows are 16-byte aligned for 32-column chunks (ldo/ldr/N multiples of 8)
  int tma_store;  // 1: 16-bit output, identity rows, no residual -> staged through smem and written by TMA
  int coalesced;  // 1: fp32 output (+residual, +scatter) -> transposed through smem, 4 full lines per warp access
  int map_mode;
  WinGeom geom;
};

__device__ __forceinline__ long long epi_out_row(const EpiParams& ep, int row) {
  if (ep.map_mode == ROWMAP_WINDOW) {
    int b = row / ep.geom.N, r = row - b * ep.geom.N;
    return static_cast<long long>(b) * ep.geom.N + win_row_to_token(ep.geom, r);
  }
  return row;
}

// Branch-free exact-erf GELU for the tensor-core epilogues.  erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, i.e. fp32 round-off level - three orders below the 16-bit output rounding), one
// MUFU.RCP + one MUFU.EX2 + ~12 FMA-pipe instructions instead of erff()'s two divergent branches.
// The fp32 validation mode keeps erff() (gelu_erf).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);   // 0.5x(1 + sign(x) erf|z|) = hx + |hx| erf_abs
}

__device__ __forceinline__ float epi_act(int act, float v) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// Scalar form (SIMT kernel and ragged edges).
__device__ __forceinline__ void epi_store_scalar(const EpiParams& ep, long long orow, int col, float acc) {
  float v = acc;
  if (ep.bias) v += __ldg(ep.bias + col);
  v = epi_act(ep.act, v);
  if (ep.resid) v += ep.resid[orow * ep.ldr + col];
  if (ep.out_dtype == DT_BF16)
    reinterpret_cast<__nv_bfloat16*>(ep.out)[orow * ep.ldo + col] = __float2bfloat16_rn(v);
  else if (ep.out_dtype == DT_F16)
    reinterpret_cast<__half*>(ep.out)[orow * ep.ldo + col] = __float2half_rn(v);
  else
    reinterpret_cast<float*>(ep.out)[orow * ep.ldo + col] = v;
}

// bias + activation on 32 consecutive full columns held in registers (16-byte aligned bias).
__device__ __forceinline__ void epi_bias_act32(const EpiParams& ep, int col0, float (&v)[32]) {
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (ep.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
  } else if (ep.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}

__device__ __forceinline__ uint32_t pack16(bool bf, float lo, float hi) { return bf ? pack_bf16x2(lo, hi) : pack_f16x2(lo, hi); }

// 32 consecutive columns of one row held in registers (the tcgen05 epilogue shape), direct global store.
__device__ __forceinline__ void epi_store_chunk32(const EpiParams& ep, long long orow, int col0, const uint32_t (&r)[32]) {
  if (ep.vec_ok && col0 + 32 <= ep.N) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    epi_bias_act32(ep, col0, v);
    if (ep.resid) {
      const float4* r4 = reinterpret_cast<const float4*>(ep.resid + orow * ep.ldr + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = r4[j];
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    if (ep.out_dtype != DT_F32) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ep.out) + orow * ep.ldo + col0);
      const bool bf = ep.out_dtype == DT_BF16;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_uint4(pack16(bf, v[8 * j + 0], v[8 * j + 1]), pack16(bf, v[8 * j + 2], v[8 * j + 3]),
                          pack16(bf, v[8 * j + 4], v[8 * j + 5]), pack16(bf, v[8 * j + 6], v[8 * j + 7]));
    } else {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow * ep.ldo + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < ep.N) epi_store_scalar(ep, orow, col0 + j, __uint_as_float(r[j]));
  }
}

// Host launchers (gemm.cu).  in_dtype: DT_BF16 / DT_F16 -> tcgen05 kind::f16 (W in the same format);
// DT_F32 -> tcgen05 kind::tf32 when impl == GEMM_TC, exact fp32 FMA when impl == GEMM_SIMT.
enum : int { GEMM_TC = 0, GEMM_SIMT = 1 };
struct GemmTuning {
  int max_ctas;   // 0 = one per SM
  int cluster;    // 0 = auto, else 1 / 2 / 4 CTAs sharing the weight tile by TMA multicast
  int tma_store;  // -1 = auto, 0 = direct stores, 1 = smem-staged TMA stores where legal
};
int launch_gemm(const void* A, long long lda, const void* W, long long ldw, int in_dtype, int M, int N, int K,
                const EpiParams& ep, int impl, const GemmTuning& tune, cudaStream_t stream);

}  // namespace csvit
