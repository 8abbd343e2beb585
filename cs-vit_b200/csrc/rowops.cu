// HBM-bound row kernels: LayerNorm with gather addressing, per-channel affine (folded BatchNorm), patch
// im2col with the image normalisation folded in, and the integer-map dump kernels used by the bit-exact tests.
//
// Reference ops replaced: K2 layernorm_before + K3 pad/roll/window_partition
// (HF:swin/modeling_swin.py:606-622), K12's layernorm_after (:648), K13's 2x2 concat + norm (:338-347),
// K14 final LayerNorm (:882), K1's Normalize + conv unfold (ref:cs_vit/net/ti_poser.py:239-243,425;
// HF:swin/modeling_swin.py:286-295), and the BatchNorm1d transposes of the head
// (ref:cs_vit/net/transformer_module.py:312,316).  All are one pass: read fp32 once, write once.
#include <cstdlib>
#include "errors.h"
#include "rowops.cuh"

namespace csvit {

template <typename T> struct Store4;
template <> struct Store4<float> {
  __device__ static void st(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
};
template <> struct Store4<__half> {
  __device__ static void st(__half* p, float a, float b, float c, float d) {
    uint2 v; v.x = pack_f16x2(a, b); v.y = pack_f16x2(c, d);
    *reinterpret_cast<uint2*>(p) = v;
  }
};
template <> struct Store4<__nv_bfloat16> {
  __device__ static void st(__nv_bfloat16* p, float a, float b, float c, float d) {
    uint2 v; v.x = pack_bf16x2(a, b); v.y = pack_bf16x2(c, d);
    *reinterpret_cast<uint2*>(p) = v;
  }
};

// One warp per R consecutive output rows; the rows (<= MAXJ*128 floats each) live in registers between the two
// passes, and all R rows' loads are issued before the first reduction so each lane keeps R*MAXJ 16-byte
// requests in flight (the kernel is purely HBM-bound).
template <int MAXJ, int R, typename OutT>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
               OutT* __restrict__ out, long long ldo, int rows, int C, int mode, WinGeom g, int scatter_min_c) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int Cout = mode == LN_MERGE2X2 ? 4 * C : C;
  const int n4 = Cout >> 2;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  // grid-stride over row groups: with a grid of (SMs x resident CTAs) every SM streams until the end instead of
  // finishing in 2-3 uneven waves of short-lived CTAs
  for (int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; row0 < rows; row0 += warps_total * R) {
  float4 v[R][MAXJ];
  float sum[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row0 + r;
    sum[r] = 0.f;
    if (row < rows) {
      long long src[4];
      if (mode == LN_IDENTITY || (mode == LN_WINDOW && C >= scatter_min_c)) {
        src[0] = row;   // wide rows in window mode: walk the SOURCE tokens in memory order (sequential fp32 reads, 2/3 of
                        // the traffic) and scatter the 16-bit rows (>= 1 KB each) to their window-order position
      } else if (mode == LN_WINDOW) {
        int b = row / g.N, rr = row - b * g.N;
        src[0] = static_cast<long long>(b) * g.N + win_row_to_token(g, rr);   // narrow rows: gather reads, dense writes
      } else {
        const int Wo = g.W >> 1, No = (g.H >> 1) * Wo;
        int b = row / No, t = row - b * No;
        int Y = t / Wo, X = t - Y * Wo;
#pragma unroll
        for (int q = 0; q < 4; ++q) src[q] = static_cast<long long>(b) * g.N + merge_src_token(g.W, Y, X, q);
      }
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        int i4 = lane + 32 * j;
        if (i4 < n4) {
          int e = i4 << 2;
          const float* p;
          if (mode == LN_MERGE2X2) { int q = e / C; p = x + src[q] * C + (e - q * C); }
          else p = x + src[0] * C + e;
          v[r][j] = *reinterpret_cast<const float4*>(p);
        }
      }
    }
  }
  // scatter destination of row row0 + r, evaluated by lane r (the map costs six integer divisions: done once per lane instead of
  // 32 times per row, it no longer dominates the instruction stream of narrow rows)
  int my_drow = 0;
  const bool scatter = mode == LN_WINDOW && C >= scatter_min_c;
  if (scatter && lane < R && row0 + lane < rows) {
    const int row = row0 + lane;
    const int b = row / g.N, t = row - b * g.N;
    my_drow = b * g.N + win_token_to_row(g, t);      // rows < 2^31 (checked by the launcher's int row count)
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = row0 + r;
    if (row >= rows) break;   // warp-uniform
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < n4) sum[r] += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
    const float mean = warp_sum(sum[r]) / float(Cout);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (lane + 32 * j < n4) {
        float a = v[r][j].x - mean, b = v[r][j].y - mean, c = v[r][j].z - mean, d = v[r][j].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / float(Cout) + eps);
    long long drow = row;
    if (scatter) drow = __shfl_sync(0xffffffffu, my_drow, r);
    OutT* orow = out + drow * ldo;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      int i4 = lane + 32 * j;
      if (i4 < n4) {
        int e = i4 << 2;
        float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + e));
        float4 bt = __ldg(reinterpret_cast<const float4*>(beta + e));
        Store4<OutT>::st(orow + e, (v[r][j].x - mean) * rstd * gm.x + bt.x, (v[r][j].y - mean) * rstd * gm.y + bt.y,
                         (v[r][j].z - mean) * rstd * gm.z + bt.z, (v[r][j].w - mean) * rstd * gm.w + bt.w);
      }
    }
  }
  }
}

// Narrow rows (C <= 256: Swin stages 0-1): LPR lanes share a row, so one instruction stream serves 32 / LPR rows at once and a
// reduction takes log2(LPR) shuffle steps.  The warp-per-row form above spends ~130 warp instructions on a 512-byte row (ncu: issue
// slots 73 % busy, 4.85 TB/s); this one ~30, which leaves the kernel to the memory system.  Each lane holds J float4 of its row
// (element i4 = lane_in_row + LPR * j: a group's LPR lanes read LPR * 16 contiguous bytes per load), gamma / beta stay in registers
// across the grid-stride loop.  Modes: identity and window-scatter (as ln_rows_kernel with C >= scatter_min_c).
template <int LPR, int J, int PASSES, typename OutT>
__global__ void __launch_bounds__(256)
ln_rows_sub_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   OutT* __restrict__ out, long long ldo, int rows, int mode, WinGeom g) {
  griddep_launch();
  griddep_wait();
  constexpr int C = LPR * J * 4;
  constexpr int RPW = 32 / LPR;                      // rows per warp per pass
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  float4 gm[J], bt[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    gm[j] = __ldg(reinterpret_cast<const float4*>(gamma) + li + LPR * j);
    bt[j] = __ldg(reinterpret_cast<const float4*>(beta) + li + LPR * j);
  }
  for (int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (RPW * PASSES); row0 < rows; row0 += warps_total * RPW * PASSES) {
    float4 v[PASSES][J];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int row = row0 + p * RPW + sub;
      if (row < rows) {
        const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
#pragma unroll
        for (int j = 0; j < J; ++j) v[p][j] = src[li + LPR * j];
      } else {
#pragma unroll
        for (int j = 0; j < J; ++j) v[p][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int row = row0 + p * RPW + sub;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) sum += (v[p][j].x + v[p][j].y) + (v[p][j].z + v[p][j].w);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.0f / float(C));
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float a = v[p][j].x - mean, b = v[p][j].y - mean, c = v[p][j].z - mean, d = v[p][j].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * (1.0f / float(C)) + eps);
      if (row < rows) {
        long long drow = row;
        if (mode == LN_WINDOW) {
          const int b = row / g.N, t = row - b * g.N;
          drow = static_cast<long long>(b) * g.N + win_token_to_row(g, t);
        }
        OutT* orow = out + drow * ldo;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int e = (li + LPR * j) << 2;
          Store4<OutT>::st(orow + e, (v[p][j].x - mean) * rstd * gm[j].x + bt[j].x, (v[p][j].y - mean) * rstd * gm[j].y + bt[j].y,
                           (v[p][j].z - mean) * rstd * gm[j].z + bt[j].z, (v[p][j].w - mean) * rstd * gm[j].w + bt[j].w);
        }
      }
    }
  }
}

template <int LPR, int J, typename OutT>
static void launch_ln_sub(const float* x, const float* gamma, const float* beta, float eps, OutT* out, long long ldo, int rows,
                          int mode, const WinGeom& g, cudaStream_t stream) {
  constexpr int PASSES = 2;
  const int rows_per_warp = (32 / LPR) * PASSES;
  const int warps = (rows + rows_per_warp - 1) / rows_per_warp;
  (void)launch_pdl(ln_rows_sub_kernel<LPR, J, PASSES, OutT>, dim3((warps + 7) / 8), dim3(256), 0, stream, x, gamma, beta, eps, out, ldo, rows, mode, g);
}

template <int MAXJ, int R, typename OutT>
static void launch_ln_cfg(const float* x, const float* gamma, const float* beta, float eps, OutT* out, long long ldo,
                          int rows, int C, int mode, const WinGeom& g, cudaStream_t stream) {
  const int warps = (rows + R - 1) / R;
  int blocks = (warps + 7) / 8;
  static int persist = -1;
  if (persist < 0) { const char* e = getenv("CSVIT_LN_PERSIST"); persist = e ? atoi(e) : 0; }
  if (persist > 0 && blocks > num_sms() * persist) blocks = num_sms() * persist;
  // window mode: rows at least this wide are read in memory order and SCATTERED as 16-bit rows; narrower rows were meant to be
  // gathered instead, but scattering measured faster down to C = 128 (tools/bench_ln.py: 221 vs 242 us at stage 0)
  static int scatter_min = -1;
  if (scatter_min < 0) { const char* e = getenv("CSVIT_LN_SCATTER_MIN_C"); scatter_min = e ? atoi(e) : 128; }
  (void)launch_pdl(ln_rows_kernel<MAXJ, R, OutT>, dim3(blocks), dim3(256), 0, stream, x, gamma, beta, eps, out, ldo, rows, C, mode, g, scatter_min);
}

template <typename OutT>
static int launch_ln_t(const float* x, const float* gamma, const float* beta, float eps, OutT* out, long long ldo,
                       int rows, int C, int mode, const WinGeom& g, cudaStream_t stream) {
  const int Cout = mode == LN_MERGE2X2 ? 4 * C : C;
  static const bool sub_rows = [] { const char* e = getenv("CSVIT_LN_SUBROWS"); return !(e && e[0] == '0'); }();
  if (sub_rows && mode != LN_MERGE2X2 && (C == 96 || C == 128 || C == 192 || C == 256)) {   // narrow rows: several rows per warp pass
    if (C == 128) launch_ln_sub<8, 4, OutT>(x, gamma, beta, eps, out, ldo, rows, mode, g, stream);
    else if (C == 96) launch_ln_sub<8, 3, OutT>(x, gamma, beta, eps, out, ldo, rows, mode, g, stream);
    else if (C == 256) launch_ln_sub<16, 4, OutT>(x, gamma, beta, eps, out, ldo, rows, mode, g, stream);
    else launch_ln_sub<16, 3, OutT>(x, gamma, beta, eps, out, ldo, rows, mode, g, stream);
    CSVIT_CUDA(cudaGetLastError());
    return 0;
  }
  // (the same layout with LPR = 32 at C = 512 measured 30.7 vs 29.0 us: wide rows are not instruction-bound, they keep the warp-per-row form)
  if (Cout <= 128) launch_ln_cfg<1, 8, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else if (Cout <= 256) launch_ln_cfg<2, 4, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else if (Cout <= 512) launch_ln_cfg<4, 2, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else if (Cout <= 1024) launch_ln_cfg<8, 2, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else if (Cout <= 2048) launch_ln_cfg<16, 1, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else if (Cout <= 4096) launch_ln_cfg<32, 1, OutT>(x, gamma, beta, eps, out, ldo, rows, C, mode, g, stream);
  else return set_error("layernorm: row width %d exceeds 4096", Cout);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, void* out, int out_dtype,
                     long long ldo, int rows, int C, int mode, const WinGeom& g, cudaStream_t stream) {
  CSVIT_REQUIRE(C % 4 == 0, "layernorm: C=%d must be a multiple of 4", C);
  CSVIT_REQUIRE((ldo % 4) == 0, "layernorm: ldo=%lld must be a multiple of 4", ldo);
  if (rows <= 0) return 0;
  if (out_dtype == DT_BF16)
    return launch_ln_t<__nv_bfloat16>(x, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, rows, C, mode, g, stream);
  if (out_dtype == DT_F16)
    return launch_ln_t<__half>(x, gamma, beta, eps, static_cast<__half*>(out), ldo, rows, C, mode, g, stream);
  return launch_ln_t<float>(x, gamma, beta, eps, static_cast<float*>(out), ldo, rows, C, mode, g, stream);
}

// y[r, c] = x[r, c] * scale[c] + shift[c] (+ add[r % add_rows, c])      (eval-mode BatchNorm1d, PE add)
template <typename OutT>
__global__ void __launch_bounds__(256)
affine_rows_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   OutT* __restrict__ out, long long total4, int C4) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total4; i += stride) {
    int c4 = static_cast<int>(i % C4);
    float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 s = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    float4 t = __ldg(reinterpret_cast<const float4*>(shift) + c4);
    Store4<OutT>::st(out + (i << 2), v.x * s.x + t.x, v.y * s.y + t.y, v.z * s.z + t.z, v.w * s.w + t.w);
  }
}

int launch_affine_rows(const float* x, const float* scale, const float* shift, void* out, int out_dtype, long long rows,
                       int C, cudaStream_t stream) {
  CSVIT_REQUIRE(C % 4 == 0, "affine_rows: C=%d must be a multiple of 4", C);
  const long long total4 = rows * (C / 4);
  if (total4 <= 0) return 0;
  int blocks = static_cast<int>((total4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_dtype == DT_BF16)
    affine_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(x, scale, shift, static_cast<__nv_bfloat16*>(out), total4, C / 4);
  else if (out_dtype == DT_F16)
    affine_rows_kernel<__half><<<blocks, 256, 0, stream>>>(x, scale, shift, static_cast<__half*>(out), total4, C / 4);
  else
    affine_rows_kernel<float><<<blocks, 256, 0, stream>>>(x, scale, shift, static_cast<float*>(out), total4, C / 4);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// Patch unfold: out[(b, py, px), c*16 + ky*4 + kx] = (img[b, c, 4py+ky, 4px+kx] - mean[c]) / std[c].
// One thread per (b, c, y, px): 16-byte coalesced reads along an image row.
template <typename OutT>
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const float* __restrict__ img, OutT* __restrict__ out, int B, int S, float m0, float m1, float m2,
                    float is0, float is1, float is2) {
  const int P = S >> 2;
  const long long total = static_cast<long long>(B) * 3 * S * P;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    int px = static_cast<int>(i % P);
    long long t = i / P;
    int y = static_cast<int>(t % S); t /= S;
    int c = static_cast<int>(t % 3);
    int b = static_cast<int>(t / 3);
    float4 v = reinterpret_cast<const float4*>(img)[i];
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float is = c == 0 ? is0 : (c == 1 ? is1 : is2);
    const int py = y >> 2, ky = y & 3;
    OutT* o = out + ((static_cast<long long>(b) * P + py) * P + px) * 48 + c * 16 + ky * 4;
    Store4<OutT>::st(o, (v.x - m) * is, (v.y - m) * is, (v.z - m) * is, (v.w - m) * is);
  }
}

int launch_patch_im2col(const float* img, void* out, int out_dtype, int B, int S, const float* mean3, const float* std3,
                        cudaStream_t stream) {
  CSVIT_REQUIRE(S % 4 == 0, "patch_im2col: image side %d must be a multiple of 4", S);
  const long long total = static_cast<long long>(B) * 3 * S * (S / 4);
  if (total <= 0) return 0;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  const float i0 = 1.0f / std3[0], i1 = 1.0f / std3[1], i2 = 1.0f / std3[2];
  if (out_dtype == DT_BF16)
    patch_im2col_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(img, static_cast<__nv_bfloat16*>(out), B, S, mean3[0], mean3[1], mean3[2], i0, i1, i2);
  else if (out_dtype == DT_F16)
    patch_im2col_kernel<__half><<<blocks, 256, 0, stream>>>(img, static_cast<__half*>(out), B, S, mean3[0], mean3[1], mean3[2], i0, i1, i2);
  else
    patch_im2col_kernel<float><<<blocks, 256, 0, stream>>>(img, static_cast<float*>(out), B, S, mean3[0], mean3[1], mean3[2], i0, i1, i2);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// ---- integer-map dumps (bit-exact tests; the hot path never materialises these) ----
__global__ void window_index_map_kernel(int* out, WinGeom g) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < g.N) out[r] = win_row_to_token(g, r);
}
__global__ void shift_mask_kernel(float* out, WinGeom g, int nW) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = nW * g.L * g.L;
  if (idx >= total) return;
  int w = idx / (g.L * g.L), rem = idx - w * g.L * g.L;
  int i = rem / g.L, j = rem - i * g.L;
  out[idx] = (g.shift > 0 && win_region(g, w, i) != win_region(g, w, j)) ? -100.0f : 0.0f;
}
__global__ void rel_index_kernel(int* out, int ws) {
  int L = ws * ws;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < L * L) out[idx] = rel_pos_index(ws, idx / L, idx % L);
}
__global__ void merge_index_map_kernel(int* out, int H, int W) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int No = (H / 2) * (W / 2);
  if (idx >= No * 4) return;
  int t = idx >> 2, q = idx & 3;
  out[idx] = merge_src_token(W, t / (W / 2), t % (W / 2), q);
}
// bias_exp[h, i, j] = table[rel_pos_index(i, j), h]   (HF:swin/modeling_swin.py:428-434)
__global__ void expand_rel_bias_kernel(const float* __restrict__ table, float* __restrict__ out, int heads, int ws) {
  int L = ws * ws;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * L * L) return;
  int h = idx / (L * L), rem = idx - h * L * L;
  out[idx] = table[rel_pos_index(ws, rem / L, rem % L) * heads + h];
}

int launch_window_index_map(int H, int W, int ws, int shift, int* out, cudaStream_t stream) {
  WinGeom g = make_geom(H, W, ws, shift);
  window_index_map_kernel<<<(g.N + 255) / 256, 256, 0, stream>>>(out, g);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}
int launch_shift_mask(int H, int W, int ws, int shift, float* out, cudaStream_t stream) {
  WinGeom g = make_geom(H, W, ws, shift);
  int nW = (H / ws) * (W / ws);
  int total = nW * g.L * g.L;
  shift_mask_kernel<<<(total + 255) / 256, 256, 0, stream>>>(out, g, nW);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}
int launch_rel_index(int ws, int* out, cudaStream_t stream) {
  int total = ws * ws * ws * ws;
  rel_index_kernel<<<(total + 255) / 256, 256, 0, stream>>>(out, ws);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}
int launch_merge_index_map(int H, int W, int* out, cudaStream_t stream) {
  int total = (H / 2) * (W / 2) * 4;
  merge_index_map_kernel<<<(total + 255) / 256, 256, 0, stream>>>(out, H, W);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}
int launch_expand_rel_bias(const float* table, float* out, int heads, int ws, cudaStream_t stream) {
  int total = heads * ws * ws * ws * ws;
  expand_rel_bias_kernel<<<(total + 255) / 256, 256, 0, stream>>>(table, out, heads, ws);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace csvit
