// Backward-pass kernels of the finetune step (BASELINE configs[3]: forward + backward of the spatial model).
// torch autograd derives these from the eager ops of HF:swin/modeling_swin.py:591-653 (SwinLayer) and
// ref:cs_vit/net/transformer_module.py:250-378; here each is one kernel on the layouts of the forward path:
//
//   col_reduce_kernel      column sums over rows (bias gradients, BatchNorm batch statistics and their gradients),
//                          optionally writing a 16-bit / window-gathered copy of the rows it reads (the fp32 residual
//                          gradient becomes a tensor-core operand in the same pass)
//   eltwise_kernel         exact-erf GELU forward / backward, ReLU backward
//   ln_bwd_kernel          LayerNorm backward with the forward's gather modes (identity / shifted-window / 2x2 merge)
//                          fused with the residual-gradient add; gamma / beta gradients reduced per CTA
//   attention_bwd_kernel   softmax attention backward for <= 64 keys, head_dim 32 (Swin windows with bias + shift mask,
//                          and the head's MHA whose logits are multiplied by sqrt(d), quirk Q1), exact fp32 math
//   affine2_rows_kernel    out = a[c] dy + b[c] x + c0[c] (+ resid): BatchNorm1d backward applied per channel
#include "errors.h"
#include "rowops.cuh"

namespace csvit {

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void st4<__half>(__half* p, float4 v) {
  uint2 u; u.x = pack_f16x2(v.x, v.y); u.y = pack_f16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ----------------------------------------------------------------------------------------------------
// Column reductions (+ optional cast / gather copy)
// ----------------------------------------------------------------------------------------------------
enum : int { CR_SUM = 0, CR_CENTERED = 1, CR_DOT = 2 };
// a: SrcT [rows, C] (pitch lda).  Row r of the pass reads source row map(r) (identity or window gather, as csvit_layernorm).
//   CR_SUM      s1[c] += sum_r a[r,c]
//   CR_CENTERED s1[c] += sum_r (a - center[c]),  s2[c] += sum_r (a - center[c])^2
//   CR_DOT      s1[c] += sum_r a[r,c],           s2[c] += sum_r a[r,c] * b[r,c]         (b fp32, pitch ldb, same row map)
// copy (optional): copy[r, c] = a[map(r), c] in OutT.   s1 / s2 may be null (pure cast / gather).
// Grid: x = row blocks of `rows_per_cta`, y = groups of 128 columns (one float4 per lane), so wide-and-short tensors (the
// [6272, 2048] hidden gradient of stage 2) still fill the machine; each warp streams rows r0 + warp, + 8, ...
template <typename SrcT, typename OutT>
__global__ void __launch_bounds__(256)
col_reduce_kernel(const SrcT* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
                  const float* __restrict__ center, int mode, int rows, int C, int row_mode, WinGeom g, OutT* __restrict__ copy,
                  long long ldc, float* __restrict__ s1, float* __restrict__ s2, int rows_per_cta) {
  __shared__ float4 red1[8][32];
  __shared__ float4 red2[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = C >> 2;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const int c4 = blockIdx.y * 32 + lane;
  const bool ok = c4 < n4;
  float4 acc1 = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc1, ctr = acc1;
  if (ok && mode == CR_CENTERED) ctr = *reinterpret_cast<const float4*>(center + 4 * c4);
  if (ok) {
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += 8) {
      long long src = r;
      if (row_mode == LN_WINDOW) {
        const int bi = r / g.N, rr = r - bi * g.N;
        src = static_cast<long long>(bi) * g.N + win_row_to_token(g, rr);
      }
      float4 v = ld4<SrcT>(a + src * lda + 4 * c4);
      if (copy) st4<OutT>(copy + static_cast<long long>(r) * ldc + 4 * c4, v);
      if (mode == CR_CENTERED) {
        v.x -= ctr.x; v.y -= ctr.y; v.z -= ctr.z; v.w -= ctr.w;
        acc2.x = fmaf(v.x, v.x, acc2.x); acc2.y = fmaf(v.y, v.y, acc2.y); acc2.z = fmaf(v.z, v.z, acc2.z); acc2.w = fmaf(v.w, v.w, acc2.w);
      } else if (mode == CR_DOT) {
        const float4 w = *reinterpret_cast<const float4*>(b + src * ldb + 4 * c4);
        acc2.x = fmaf(v.x, w.x, acc2.x); acc2.y = fmaf(v.y, w.y, acc2.y); acc2.z = fmaf(v.z, w.z, acc2.z); acc2.w = fmaf(v.w, w.w, acc2.w);
      }
      acc1.x += v.x; acc1.y += v.y; acc1.z += v.z; acc1.w += v.w;
    }
  }
  if (s1 == nullptr && s2 == nullptr) return;
  red1[warp][lane] = acc1;
  red2[warp][lane] = acc2;
  __syncthreads();
  if (warp == 0 && ok) {
    float4 t1 = red1[0][lane], t2 = red2[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 u1 = red1[w][lane], u2 = red2[w][lane];
      t1.x += u1.x; t1.y += u1.y; t1.z += u1.z; t1.w += u1.w;
      t2.x += u2.x; t2.y += u2.y; t2.z += u2.z; t2.w += u2.w;
    }
    if (s1) { atomicAdd(s1 + 4 * c4, t1.x); atomicAdd(s1 + 4 * c4 + 1, t1.y); atomicAdd(s1 + 4 * c4 + 2, t1.z); atomicAdd(s1 + 4 * c4 + 3, t1.w); }
    if (s2 && mode != CR_SUM) { atomicAdd(s2 + 4 * c4, t2.x); atomicAdd(s2 + 4 * c4 + 1, t2.y); atomicAdd(s2 + 4 * c4 + 2, t2.z); atomicAdd(s2 + 4 * c4 + 3, t2.w); }
  }
}

template <typename SrcT, typename OutT>
static int launch_cr_t(const void* a, long long lda, const float* b, long long ldb, const float* center, int mode, int rows, int C,
                       int row_mode, const WinGeom& g, void* copy, long long ldc, float* s1, float* s2, cudaStream_t stream) {
  const int col_groups = (C / 4 + 31) / 32;
  // rows per CTA: as many as keep about 4 CTAs per SM in flight, between 64 and 512 (fewer atomics, longer streams)
  int rpc = 512;
  while (rpc > 64 && static_cast<long long>((rows + rpc - 1) / rpc) * col_groups < 4ll * 148) rpc >>= 1;
  dim3 grid((rows + rpc - 1) / rpc, col_groups);
  CSVIT_REQUIRE(grid.y < 65536, "col_reduce: too many column groups");
  col_reduce_kernel<SrcT, OutT><<<grid, 256, 0, stream>>>(static_cast<const SrcT*>(a), lda, b, ldb, center, mode, rows, C, row_mode, g,
                                                          static_cast<OutT*>(copy), ldc, s1, s2, rpc);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

template <typename SrcT>
static int launch_cr_s(const void* a, long long lda, const float* b, long long ldb, const float* center, int mode, int rows, int C,
                       int row_mode, const WinGeom& g, void* copy, int copy_dtype, long long ldc, float* s1, float* s2,
                       cudaStream_t stream) {
  if (copy_dtype == DT_BF16) return launch_cr_t<SrcT, __nv_bfloat16>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, ldc, s1, s2, stream);
  if (copy_dtype == DT_F16) return launch_cr_t<SrcT, __half>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, ldc, s1, s2, stream);
  return launch_cr_t<SrcT, float>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, ldc, s1, s2, stream);
}

int launch_col_reduce(const void* a, int a_dtype, long long lda, const float* b, long long ldb, const float* center, int mode,
                      int rows, int C, int row_mode, const WinGeom& g, void* copy, int copy_dtype, long long ldc, float* s1,
                      float* s2, cudaStream_t stream) {
  CSVIT_REQUIRE(C % 4 == 0, "col_reduce: C=%d must be a multiple of 4", C);
  CSVIT_REQUIRE(mode >= CR_SUM && mode <= CR_DOT, "col_reduce: bad mode %d", mode);
  CSVIT_REQUIRE(mode != CR_CENTERED || center != nullptr, "col_reduce: centered mode needs the centre vector");
  CSVIT_REQUIRE(mode != CR_DOT || b != nullptr, "col_reduce: dot mode needs the second operand");
  CSVIT_REQUIRE(row_mode == LN_IDENTITY || row_mode == LN_WINDOW, "col_reduce: bad row mode %d", row_mode);
  if (rows <= 0) return 0;
  if (a_dtype == DT_BF16) return launch_cr_s<__nv_bfloat16>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, copy_dtype, ldc, s1, s2, stream);
  if (a_dtype == DT_F16) return launch_cr_s<__half>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, copy_dtype, ldc, s1, s2, stream);
  return launch_cr_s<float>(a, lda, b, ldb, center, mode, rows, C, row_mode, g, copy, copy_dtype, ldc, s1, s2, stream);
}

// ----------------------------------------------------------------------------------------------------
// Elementwise: GELU forward / backward (exact erf, nn.GELU()), ReLU backward
// ----------------------------------------------------------------------------------------------------
enum : int { EW_GELU_FWD = 0, EW_GELU_BWD = 1, EW_RELU_BWD = 2 };

__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

// GELU_FWD: out = gelu(a).  GELU_BWD: out = a * gelu'(b)  (a = dy, b = pre-activation).  RELU_BWD: out = b > 0 ? a : 0 (b = output).
template <typename T>
__global__ void __launch_bounds__(256)
eltwise_kernel(int op, const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    float4 x = ld4<T>(a + 4 * i), r;
    if (op == EW_GELU_FWD) {
      r = make_float4(gelu_erf(x.x), gelu_erf(x.y), gelu_erf(x.z), gelu_erf(x.w));
    } else {
      const float4 y = ld4<T>(b + 4 * i);
      if (op == EW_GELU_BWD) r = make_float4(x.x * gelu_grad(y.x), x.y * gelu_grad(y.y), x.z * gelu_grad(y.z), x.w * gelu_grad(y.w));
      else r = make_float4(y.x > 0.f ? x.x : 0.f, y.y > 0.f ? x.y : 0.f, y.z > 0.f ? x.z : 0.f, y.w > 0.f ? x.w : 0.f);
    }
    st4<T>(out + 4 * i, r);
  }
}

// out[r, :] = (x ? x[r, :] : 0) + s[r / group_rows] * y[r, :]   (fp32, C a multiple of 4): the per-sample stochastic-depth
// scale of the attention branch, forward (x = shortcut) and backward (x = nullptr: the branch's share of the gradient).
__global__ void __launch_bounds__(256)
row_scale_add_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float* __restrict__ s, float4* __restrict__ out,
                     long long n4, int c4, int group_rows) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float k = s[(i / c4) / group_rows];
    const float4 b = y[i];
    float4 a = x ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    a.x = fmaf(k, b.x, a.x); a.y = fmaf(k, b.y, a.y); a.z = fmaf(k, b.z, a.z); a.w = fmaf(k, b.w, a.w);
    out[i] = a;
  }
}

int launch_row_scale_add(const float* x, const float* y, const float* s, float* out, long long rows, int C, int group_rows,
                         cudaStream_t stream) {
  CSVIT_REQUIRE(C > 0 && C % 4 == 0 && group_rows > 0 && rows % group_rows == 0, "row_scale_add: C=%d rows=%lld group=%d", C, rows, group_rows);
  const long long n4 = rows * (C / 4);
  if (n4 == 0) return 0;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148ll * 16) blocks = 148ll * 16;
  row_scale_add_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), s,
                                                                     reinterpret_cast<float4*>(out), n4, C / 4, group_rows);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_eltwise(int op, const void* a, const void* b, void* out, int dtype, long long n, cudaStream_t stream) {
  CSVIT_REQUIRE(op >= EW_GELU_FWD && op <= EW_RELU_BWD, "eltwise: bad op %d", op);
  CSVIT_REQUIRE(n % 4 == 0, "eltwise: element count %lld must be a multiple of 4", n);
  CSVIT_REQUIRE(op == EW_GELU_FWD || b != nullptr, "eltwise: backward ops need the second operand");
  if (n <= 0) return 0;
  const long long n4 = n / 4;
  const int blocks = int(n4 / 256 + 1 < 148 * 16 ? n4 / 256 + 1 : 148 * 16);
  if (dtype == DT_BF16)
    eltwise_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(op, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
                                                              static_cast<__nv_bfloat16*>(out), n4);
  else if (dtype == DT_F16)
    eltwise_kernel<__half><<<blocks, 256, 0, stream>>>(op, static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(out), n4);
  else
    eltwise_kernel<float><<<blocks, 256, 0, stream>>>(op, static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(out), n4);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// out[r,c] = a[c] * dy[r,c] + b[c] * x[r,c] + c0[c] (+ resid[r,c]); all fp32, dense rows.
__global__ void __launch_bounds__(256)
affine2_rows_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                    const float* __restrict__ c0, const float* __restrict__ resid, float* __restrict__ out, long long rows, int C) {
  const int n4 = C >> 2;
  const long long total = rows * n4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c4 = int(i % n4);
    const float4 d = *reinterpret_cast<const float4*>(dy + 4 * i), v = *reinterpret_cast<const float4*>(x + 4 * i);
    const float4 aa = *reinterpret_cast<const float4*>(a + 4 * c4), bb = *reinterpret_cast<const float4*>(b + 4 * c4);
    const float4 cc = *reinterpret_cast<const float4*>(c0 + 4 * c4);
    float4 r = make_float4(fmaf(aa.x, d.x, fmaf(bb.x, v.x, cc.x)), fmaf(aa.y, d.y, fmaf(bb.y, v.y, cc.y)),
                           fmaf(aa.z, d.z, fmaf(bb.z, v.z, cc.z)), fmaf(aa.w, d.w, fmaf(bb.w, v.w, cc.w)));
    if (resid) {
      const float4 q = *reinterpret_cast<const float4*>(resid + 4 * i);
      r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
    }
    *reinterpret_cast<float4*>(out + 4 * i) = r;
  }
}

int launch_affine2_rows(const float* dy, const float* x, const float* a, const float* b, const float* c0, const float* resid,
                        float* out, long long rows, int C, cudaStream_t stream) {
  CSVIT_REQUIRE(C % 4 == 0, "affine2_rows: C=%d must be a multiple of 4", C);
  if (rows <= 0) return 0;
  const long long total = rows * (C / 4);
  const int blocks = int(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  affine2_rows_kernel<<<blocks, 256, 0, stream>>>(dy, x, a, b, c0, resid, out, rows, C);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// dst[c, r] = src[r, c] for fp32 matrices (32x32 tiles through padded shared memory).  The tcgen05 kind::tf32 path takes
// K-major operands only here (MN-major tf32 needs the 32-byte-atom swizzle), so the fp32 head transposes its small
// backward operands instead.
__global__ void __launch_bounds__(256)
transpose_f32_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, int R, int C) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < C) tile[i][tx] = src[static_cast<long long>(r0 + i) * lds + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && r0 + tx < R) dst[static_cast<long long>(c0 + i) * ldd + r0 + tx] = tile[tx][i];
}

int launch_transpose_f32(const float* src, long long lds, float* dst, long long ldd, int R, int C, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  CSVIT_REQUIRE(grid.y < 65536, "transpose: too many rows (%d)", R);
  transpose_f32_kernel<<<grid, 256, 0, stream>>>(src, lds, dst, ldd, R, C);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------------------------------
// LayerNorm backward
// ----------------------------------------------------------------------------------------------------
// Output row r of the forward (dy row r, width Cout) was LayerNorm of source row(s) map(r) of x (see ln_rows_kernel):
//   dx[map(r)] = (dres ? dres[map(r)] : 0) + rstd * (g - mean(g) - xhat * mean(g * xhat)),   g = dy * gamma
//   dgamma += sum_r dy * xhat,  dbeta += sum_r dy                (accumulated into fp32 vectors with atomics)
// One warp per row, persistent over rows; the row statistics are recomputed from x (fp32, two passes in registers).
template <int MAXJ, typename DyT>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const DyT* __restrict__ dy, long long ldy, const float* __restrict__ gamma, float eps,
              int rows, int C, int mode, WinGeom g, const float* __restrict__ dres, float* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sacc[];   // [2][Cout]
  const int lane = threadIdx.x & 31;
  const int Cout = mode == LN_MERGE2X2 ? 4 * C : C;
  const int n4 = Cout >> 2;
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps_total) {
    long long src[4];
    if (mode == LN_IDENTITY) {
      src[0] = row;
    } else if (mode == LN_WINDOW) {
      const int b = row / g.N, rr = row - b * g.N;
      src[0] = static_cast<long long>(b) * g.N + win_row_to_token(g, rr);
    } else {
      const int Wo = g.W >> 1, No = (g.H >> 1) * Wo;
      const int b = row / No, t = row - b * No;
      const int Y = t / Wo, X = t - Y * Wo;
#pragma unroll
      for (int q = 0; q < 4; ++q) src[q] = static_cast<long long>(b) * g.N + merge_src_token(g.W, Y, X, q);
    }
    float4 v[MAXJ], d[MAXJ];
    long long off[MAXJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int i4 = lane + 32 * j;
      if (i4 < n4) {
        const int e = i4 << 2;
        if (mode == LN_MERGE2X2) { const int q = e / C; off[j] = src[q] * C + (e - q * C); }
        else off[j] = src[0] * C + e;
        v[j] = *reinterpret_cast<const float4*>(x + off[j]);
        d[j] = ld4<DyT>(dy + static_cast<long long>(row) * ldy + e);
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(sum) / float(Cout);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < n4) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        sq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    const float rstd = rsqrtf(warp_sum(sq) / float(Cout) + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int i4 = lane + 32 * j;
      if (i4 < n4) {
        const int e = i4 << 2;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + e));
        v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;   // xhat
        atomicAdd(&sacc[e + 0], d[j].x * v[j].x); atomicAdd(&sacc[e + 1], d[j].y * v[j].y);
        atomicAdd(&sacc[e + 2], d[j].z * v[j].z); atomicAdd(&sacc[e + 3], d[j].w * v[j].w);
        atomicAdd(&sacc[Cout + e + 0], d[j].x); atomicAdd(&sacc[Cout + e + 1], d[j].y);
        atomicAdd(&sacc[Cout + e + 2], d[j].z); atomicAdd(&sacc[Cout + e + 3], d[j].w);
        d[j].x *= gm.x; d[j].y *= gm.y; d[j].z *= gm.z; d[j].w *= gm.w;   // g = dy * gamma
        c1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
        c2 += (d[j].x * v[j].x + d[j].y * v[j].y) + (d[j].z * v[j].z + d[j].w * v[j].w);
      }
    }
    c1 = warp_sum(c1) / float(Cout);
    c2 = warp_sum(c2) / float(Cout);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < n4) {
        float4 r = make_float4(rstd * (d[j].x - c1 - v[j].x * c2), rstd * (d[j].y - c1 - v[j].y * c2),
                               rstd * (d[j].z - c1 - v[j].z * c2), rstd * (d[j].w - c1 - v[j].w * c2));
        if (dres) {
          const float4 q = *reinterpret_cast<const float4*>(dres + off[j]);
          r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
        }
        *reinterpret_cast<float4*>(dx + off[j]) = r;
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
    atomicAdd(dgamma + i, sacc[i]);
    atomicAdd(dbeta + i, sacc[Cout + i]);
  }
}

template <int MAXJ, typename DyT>
static int launch_lnb_cfg(const float* x, const void* dy, long long ldy, const float* gamma, float eps, int rows, int C, int mode,
                          const WinGeom& g, const float* dres, float* dx, float* dgamma, float* dbeta, cudaStream_t stream) {
  const int Cout = mode == LN_MERGE2X2 ? 4 * C : C;
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  ln_bwd_kernel<MAXJ, DyT><<<blocks, 256, 2 * Cout * sizeof(float), stream>>>(x, static_cast<const DyT*>(dy), ldy, gamma, eps, rows, C, mode, g,
                                                                              dres, dx, dgamma, dbeta);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

template <typename DyT>
static int launch_lnb_t(const float* x, const void* dy, long long ldy, const float* gamma, float eps, int rows, int C, int mode,
                        const WinGeom& g, const float* dres, float* dx, float* dgamma, float* dbeta, cudaStream_t s) {
  const int Cout = mode == LN_MERGE2X2 ? 4 * C : C;
  if (Cout <= 128) return launch_lnb_cfg<1, DyT>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, s);
  if (Cout <= 256) return launch_lnb_cfg<2, DyT>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, s);
  if (Cout <= 512) return launch_lnb_cfg<4, DyT>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, s);
  if (Cout <= 1024) return launch_lnb_cfg<8, DyT>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, s);
  if (Cout <= 2048) return launch_lnb_cfg<16, DyT>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, s);
  return set_error("layernorm_bwd: row width %d exceeds 2048", Cout);
}

int launch_layernorm_bwd(const float* x, const void* dy, int dy_dtype, long long ldy, const float* gamma, float eps, int rows, int C,
                         int mode, const WinGeom& g, const float* dres, float* dx, float* dgamma, float* dbeta, cudaStream_t stream) {
  CSVIT_REQUIRE(C % 4 == 0, "layernorm_bwd: C=%d must be a multiple of 4", C);
  if (rows <= 0) return 0;
  if (dy_dtype == DT_BF16) return launch_lnb_t<__nv_bfloat16>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, stream);
  if (dy_dtype == DT_F16) return launch_lnb_t<__half>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, stream);
  return launch_lnb_t<float>(x, dy, ldy, gamma, eps, rows, C, mode, g, dres, dx, dgamma, dbeta, stream);
}

// ----------------------------------------------------------------------------------------------------
// Attention backward (<= 64 keys with a bias gradient: window attention; <= 128 keys without: the heads; head_dim 32)
// ----------------------------------------------------------------------------------------------------
constexpr int AB_MAX = 128;       // longest sequence (MAXL = 128 instantiation: no bias-gradient accumulator, it would not fit)
constexpr int AB_HD = 32;
constexpr int AB_LD = 36;         // padded fp32 row of a 32-wide operand: 16-byte aligned rows, conflict-free LDS.128
constexpr int AB_THREADS = 256;
template <int MAXL>
constexpr size_t ab_smem() {
  return sizeof(float) * (4 * MAXL * AB_LD + 2 * MAXL * (MAXL + 4) + (MAXL <= 64 ? MAXL * MAXL : 0) + 2 * MAXL) + sizeof(int) * MAXL;
}

// Forward (attention_simt_kernel): s_ij = scale q_i.k_j + bias[h,i,j] + mask_ij, p = softmax_j(s), o_i = sum_j p_ij v_j.
// Backward: dp_ij = do_i.v_j, D_i = sum_j p_ij dp_ij, ds_ij = p_ij (dp_ij - D_i),
//           dq_i = scale sum_j ds_ij k_j,  dk_j = scale sum_i ds_ij q_i,  dv_j = sum_i p_ij do_i,  dbias[h,i,j] += ds_ij.
// CTA b works on head (b % heads) only, so the bias gradient accumulates in shared memory and reaches HBM once per CTA.
template <typename T, int MAXL>
__global__ void __launch_bounds__(AB_THREADS)
attention_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ dout,
                     T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv, long long ldq, long long ldk, long long ldv,
                     long long ldo, long long lddq, long long lddk, long long lddv, int n_seq, int Lq, int S, int heads, float scale,
                     const float* __restrict__ bias, float* __restrict__ dbias, WinGeom g, int nW) {
  constexpr int AB_LDP = MAXL + 4, NH = MAXL / 32;
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                          // [MAXL][AB_LD]
  float* Ks = Qs + MAXL * AB_LD;
  float* Vs = Ks + MAXL * AB_LD;
  float* Os = Vs + MAXL * AB_LD;           // dO
  float* Ps = Os + MAXL * AB_LD;           // [MAXL][AB_LDP]
  float* Ds = Ps + MAXL * AB_LDP;          // dS
  float* Bacc = Ds + MAXL * AB_LDP;        // [Lq][S] bias-gradient accumulator (MAXL <= 64 only)
  int* region = reinterpret_cast<int*>(Bacc + (MAXL <= 64 ? MAXL * MAXL : 0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % heads;
  const int seq0 = blockIdx.x / heads, seq_step = gridDim.x / heads;
  if (MAXL <= 64 && dbias) for (int i = tid; i < Lq * S; i += AB_THREADS) Bacc[i] = 0.f;
  for (int seq = seq0; seq < n_seq; seq += seq_step) {
    __syncthreads();
    for (int idx = tid; idx < S * (AB_HD / 4); idx += AB_THREADS) {
      const int r = idx >> 3, d4 = (idx & 7) << 2;
      *reinterpret_cast<float4*>(Ks + r * AB_LD + d4) = ld4<T>(k + (static_cast<long long>(seq) * S + r) * ldk + h * AB_HD + d4);
      *reinterpret_cast<float4*>(Vs + r * AB_LD + d4) = ld4<T>(v + (static_cast<long long>(seq) * S + r) * ldv + h * AB_HD + d4);
    }
    for (int idx = tid; idx < Lq * (AB_HD / 4); idx += AB_THREADS) {
      const int r = idx >> 3, d4 = (idx & 7) << 2;
      *reinterpret_cast<float4*>(Qs + r * AB_LD + d4) = ld4<T>(q + (static_cast<long long>(seq) * Lq + r) * ldq + h * AB_HD + d4);
      *reinterpret_cast<float4*>(Os + r * AB_LD + d4) = ld4<T>(dout + (static_cast<long long>(seq) * Lq + r) * ldo + h * AB_HD + d4);
    }
    if (tid < MAXL) region[tid] = (g.shift > 0 && tid < S) ? win_region(g, seq % nW, tid) : 0;
    __syncthreads();
    // phase 1: one warp per query row, lane = key (NH keys per lane)
    for (int i = warp; i < Lq; i += AB_THREADS / 32) {
      float sc[NH], dp[NH];
#pragma unroll
      for (int half = 0; half < NH; ++half) {
        const int j = lane + 32 * half;
        float a = -INFINITY, b = 0.f;
        if (j < S) {
          a = 0.f;
#pragma unroll
          for (int d4 = 0; d4 < AB_HD; d4 += 4) {
            const float4 qq = *reinterpret_cast<const float4*>(Qs + i * AB_LD + d4);
            const float4 oo = *reinterpret_cast<const float4*>(Os + i * AB_LD + d4);
            const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AB_LD + d4);
            const float4 vv = *reinterpret_cast<const float4*>(Vs + j * AB_LD + d4);
            a = fmaf(qq.x, kk.x, a); a = fmaf(qq.y, kk.y, a); a = fmaf(qq.z, kk.z, a); a = fmaf(qq.w, kk.w, a);
            b = fmaf(oo.x, vv.x, b); b = fmaf(oo.y, vv.y, b); b = fmaf(oo.z, vv.z, b); b = fmaf(oo.w, vv.w, b);
          }
          a *= scale;
          if (bias) a += __ldg(bias + (static_cast<long long>(h) * Lq + i) * S + j);
          if (g.shift > 0 && region[j] != region[i]) a += -100.0f;
        }
        sc[half] = a; dp[half] = b;
      }
      float mloc = sc[0];
#pragma unroll
      for (int half = 1; half < NH; ++half) mloc = fmaxf(mloc, sc[half]);
      const float mx = warp_max(mloc);
      float e[NH], esum = 0.f;
#pragma unroll
      for (int half = 0; half < NH; ++half) { e[half] = lane + 32 * half < S ? expf(sc[half] - mx) : 0.f; esum += e[half]; }
      const float inv = 1.0f / warp_sum(esum);
      float dloc = 0.f;
#pragma unroll
      for (int half = 0; half < NH; ++half) { e[half] *= inv; dloc = fmaf(e[half], dp[half], dloc); }
      const float D = warp_sum(dloc);
#pragma unroll
      for (int half = 0; half < NH; ++half) {
        const float ds = e[half] * (dp[half] - D);
        Ps[i * AB_LDP + lane + 32 * half] = e[half];
        Ds[i * AB_LDP + lane + 32 * half] = ds;
        if (MAXL <= 64 && dbias && lane + 32 * half < S) Bacc[i * S + lane + 32 * half] += ds;
      }
    }
    __syncthreads();
    // phase 2: one thread per (row, 4 head dims)
    for (int idx = tid; idx < Lq * (AB_HD / 4); idx += AB_THREADS) {        // dq
      const int i = idx >> 3, d4 = (idx & 7) << 2;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < S; ++j) {
        const float w = Ds[i * AB_LDP + j];
        const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AB_LD + d4);
        acc.x = fmaf(w, kk.x, acc.x); acc.y = fmaf(w, kk.y, acc.y); acc.z = fmaf(w, kk.z, acc.z); acc.w = fmaf(w, kk.w, acc.w);
      }
      acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
      st4<T>(dq + (static_cast<long long>(seq) * Lq + i) * lddq + h * AB_HD + d4, acc);
    }
    for (int idx = tid; idx < S * (AB_HD / 4); idx += AB_THREADS) {         // dk, dv
      const int j = idx >> 3, d4 = (idx & 7) << 2;
      float4 ak = make_float4(0.f, 0.f, 0.f, 0.f), av = ak;
      for (int i = 0; i < Lq; ++i) {
        const float w = Ds[i * AB_LDP + j], p = Ps[i * AB_LDP + j];
        const float4 qq = *reinterpret_cast<const float4*>(Qs + i * AB_LD + d4);
        const float4 oo = *reinterpret_cast<const float4*>(Os + i * AB_LD + d4);
        ak.x = fmaf(w, qq.x, ak.x); ak.y = fmaf(w, qq.y, ak.y); ak.z = fmaf(w, qq.z, ak.z); ak.w = fmaf(w, qq.w, ak.w);
        av.x = fmaf(p, oo.x, av.x); av.y = fmaf(p, oo.y, av.y); av.z = fmaf(p, oo.z, av.z); av.w = fmaf(p, oo.w, av.w);
      }
      ak.x *= scale; ak.y *= scale; ak.z *= scale; ak.w *= scale;
      st4<T>(dk + (static_cast<long long>(seq) * S + j) * lddk + h * AB_HD + d4, ak);
      st4<T>(dv + (static_cast<long long>(seq) * S + j) * lddv + h * AB_HD + d4, av);
    }
  }
  __syncthreads();
  if (MAXL <= 64 && dbias)
    for (int i = tid; i < Lq * S; i += AB_THREADS) atomicAdd(dbias + static_cast<long long>(h) * Lq * S + i, Bacc[i]);
}

template <typename T, int MAXL>
static int launch_ab_t(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, long long ldq,
                       long long ldk, long long ldv, long long ldo, long long lddq, long long lddk, long long lddv, int n_seq, int Lq,
                       int S, int heads, float scale, const float* bias, float* dbias, const WinGeom& g, int nW, cudaStream_t stream) {
  static DeviceOnce once;
  auto kern = attention_bwd_kernel<T, MAXL>;
  constexpr size_t AB_SMEM = ab_smem<MAXL>();
  if (once.first()) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AB_SMEM)));
  }
  int per_head = (148 * 2 + heads - 1) / heads;
  if (per_head > n_seq) per_head = n_seq;
  if (per_head < 1) per_head = 1;
  kern<<<per_head * heads, AB_THREADS, AB_SMEM, stream>>>(
      static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<const T*>(dout), static_cast<T*>(dq),
      static_cast<T*>(dk), static_cast<T*>(dv), ldq, ldk, ldv, ldo, lddq, lddk, lddv, n_seq, Lq, S, heads, scale, bias, dbias, g, nW);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

int launch_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int dtype,
                         long long ldq, long long ldk, long long ldv, long long ldo, long long lddq, long long lddk, long long lddv,
                         int n_seq, int Lq, int S, int heads, float scale, const float* bias, float* dbias, int mH, int mW, int mws,
                         int mshift, cudaStream_t stream) {
  CSVIT_REQUIRE(S >= 1 && S <= AB_MAX && Lq >= 1 && Lq <= AB_MAX, "attention_bwd: lengths (%d, %d) outside [1,%d]", Lq, S, AB_MAX);
  CSVIT_REQUIRE(heads >= 1, "attention_bwd: heads=%d", heads);
  if (n_seq <= 0) return 0;
  WinGeom g = make_geom(mH > 0 ? mH : 1, mW > 0 ? mW : 1, mws > 0 ? mws : 1, mshift);
  const int nW = mshift > 0 ? (mH / mws) * (mW / mws) : 1;
  if (mshift > 0) CSVIT_REQUIRE(S == mws * mws && Lq == S, "attention_bwd: window mask needs Lq == S == ws^2");
  const bool wide = S > 64 || Lq > 64;
  CSVIT_REQUIRE(!wide || dbias == nullptr, "attention_bwd: a bias gradient is built for sequences of <= 64 keys only (%d, %d)", Lq, S);
#define CSVIT_AB_ARGS q, k, v, dout, dq, dk, dv, ldq, ldk, ldv, ldo, lddq, lddk, lddv, n_seq, Lq, S, heads, scale, bias, dbias, g, nW, stream
  if (dtype == DT_BF16) return wide ? launch_ab_t<__nv_bfloat16, 128>(CSVIT_AB_ARGS) : launch_ab_t<__nv_bfloat16, 64>(CSVIT_AB_ARGS);
  if (dtype == DT_F16) return wide ? launch_ab_t<__half, 128>(CSVIT_AB_ARGS) : launch_ab_t<__half, 64>(CSVIT_AB_ARGS);
  return wide ? launch_ab_t<float, 128>(CSVIT_AB_ARGS) : launch_ab_t<float, 64>(CSVIT_AB_ARGS);
#undef CSVIT_AB_ARGS
}

}  // namespace csvit
