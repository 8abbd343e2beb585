// Backward of the Swin window-attention core on tensor cores (mma.sync m16n8k16, 16-bit operands, fp32 accumulate and
// softmax), the training-path counterpart of attention.cu.  Replaces autograd's backward of HF:swin/modeling_swin.py:424-452
// (two bmm backward pairs, softmax backward, the bias-gather backward `index_put(accumulate)`).
//
// One warp per (window, head), one head per CTA - same decomposition, shared-memory layout and fragment plumbing as the
// forward kernel.  Flash-attention-style recomputation in two passes, nothing [49,49]-sized ever leaves the SM:
//   pass A (16-query-row tiles)   S = scale QK^T + bias + mask,  P = softmax(S),  dP = dO V^T,  D_i = sum_j P dP,
//                                 dS = P (dP - D_i),  dQ = scale dS K,  dbias += dS;  keeps L_i = logsumexp_i and D_i
//   pass B (16-key-row tiles)     S^T = scale K Q^T + bias^T + mask^T,  P^T = exp(S^T - L_i),  dP^T = V dO^T,
//                                 dS^T = P^T (dP^T - D_i),  dK = scale dS^T Q,  dV = P^T dO
// The transposed tiles are recomputed rather than transposed through memory: P and dS feed the second MMA of each pass
// directly from the accumulator registers (as P does in the forward kernel).
#include <type_traits>

#include "errors.h"
#include "rowops.cuh"

namespace csvit {

constexpr int BA_L = 49;
constexpr int BA_ROW = 64;                                         // bytes per 32-element 16-bit row
constexpr int BA_MAT = BA_L * BA_ROW;                              // one 49-row operand
constexpr int BA_WARP_BYTES = 4 * BA_MAT + 15 * BA_ROW + 2 * 64 * 4 + 64;   // Q K V dO + zero rows + L + D + region ids
constexpr int BA_WARPS = 4;
constexpr int BA_TABLE_BYTES = ((BA_L * BA_L * 4 + 15) / 16) * 16;

__device__ __forceinline__ uint32_t ba_off(int row, int chunk) { return uint32_t(row * BA_ROW + ((chunk ^ ((row >> 1) & 3)) << 4)); }
__device__ __forceinline__ void ba_ldsm(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ba_ldsm_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
template <typename T>
__device__ __forceinline__ void ba_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
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
__device__ __forceinline__ void ba_cp16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ba_cp_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// acc[7][4] (16 rows x 56 cols) = A_tile[16 x 32] * B[56 x 32]^T : A rows at Ab + rows m0.., B rows at Bb (both 64-byte rows).
template <typename T>
__device__ __forceinline__ void ba_tile_abt(float (&acc)[7][4], const uint8_t* Ab, int m0, const uint8_t* Bb, int lane) {
  uint32_t a[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) ba_ldsm(a[ks], Ab + ba_off(m0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    uint32_t b[4];
    ba_ldsm(b, Bb + ba_off(j * 8 + (lane & 7), lane >> 3));
    ba_mma<T>(acc[j], a[0], b[0], b[1]);
    ba_mma<T>(acc[j], a[1], b[2], b[3]);
  }
}

// out[4][4] (16 rows x 32 cols) = W[16 x 56 (padded to 64)] * B[64 x 32], W given as accumulator-layout registers.
template <typename T>
__device__ __forceinline__ void ba_tile_wb(float (&o)[4][4], const float (&w)[7][4], const uint8_t* Bb, int lane) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = Half16<T>::pack(w[2 * kk][0], w[2 * kk][1]);
    pa[1] = Half16<T>::pack(w[2 * kk][2], w[2 * kk][3]);
    if (2 * kk + 1 < 7) {
      pa[2] = Half16<T>::pack(w[2 * kk + 1][0], w[2 * kk + 1][1]);
      pa[3] = Half16<T>::pack(w[2 * kk + 1][2], w[2 * kk + 1][3]);
    } else {
      pa[2] = 0u; pa[3] = 0u;
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t vb[4];
      ba_ldsm_t(vb, Bb + ba_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)));
      ba_mma<T>(o[2 * np], pa, vb[0], vb[1]);
      ba_mma<T>(o[2 * np + 1], pa, vb[2], vb[3]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void ba_store_tile(T* dst, long long ld, long long row0, int r0, int r1, int col0, const float (&o)[4][4],
                                              float mul, int lane) {
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int col = col0 + n * 8 + (lane & 3) * 2;
    if (r0 < BA_L) *reinterpret_cast<uint32_t*>(dst + (row0 + r0) * ld + col) = Half16<T>::pack(o[n][0] * mul, o[n][1] * mul);
    if (r1 < BA_L) *reinterpret_cast<uint32_t*>(dst + (row0 + r1) * ld + col) = Half16<T>::pack(o[n][2] * mul, o[n][3] * mul);
  }
}

// qkv [B*N, 3C] window-ordered (Q | K | V), dout [B*N, C], dqkv [B*N, 3C]; bias / dbias plain [heads, 49, 49] fp32.
template <typename T>
__global__ void __launch_bounds__(BA_WARPS * 32, 3)
win_attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, T* __restrict__ dqkv, const float* __restrict__ bias,
                    float* __restrict__ dbias, int num_windows, int C, int heads, WinGeom g, int nW, float scale) {
  extern __shared__ __align__(16) uint8_t ba_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* bias_s = reinterpret_cast<float*>(ba_smem);
  float* dbias_s = reinterpret_cast<float*>(ba_smem + BA_TABLE_BYTES);
  uint8_t* Qs = ba_smem + 2 * BA_TABLE_BYTES + warp * BA_WARP_BYTES;
  uint8_t* Ks = Qs + BA_MAT;
  uint8_t* Vs = Ks + BA_MAT;
  uint8_t* Os = Vs + BA_MAT;                                   // dO, followed by 15 zero rows
  float* rowL = reinterpret_cast<float*>(Os + BA_MAT + 15 * BA_ROW);
  float* rowD = rowL + 64;
  int8_t* region_s = reinterpret_cast<int8_t*>(rowD + 64);

  const int h = blockIdx.x % heads;
  for (int i = threadIdx.x; i < BA_L * BA_L; i += BA_WARPS * 32) {
    bias_s[i] = __ldg(bias + static_cast<long long>(h) * BA_L * BA_L + i);
    dbias_s[i] = 0.f;
  }
  for (int i = lane; i < (BA_WARP_BYTES >> 2); i += 32) reinterpret_cast<uint32_t*>(Qs)[i] = 0u;
  __syncthreads();

  const int ld_qkv = 3 * C;
  const int nWy = g.H / g.ws;
  const int wstride = (gridDim.x / heads) * BA_WARPS;
  const int qc = (lane & 3) * 2;                                // this lane's column pair inside an 8-wide tile
  for (int wg = (blockIdx.x / heads) * BA_WARPS + warp; wg < num_windows; wg += wstride) {
    const int w = wg % nW;
    const long long row0 = static_cast<long long>(wg) * BA_L;
    const T* src = qkv + row0 * ld_qkv + h * 32;
    const T* dsrc = dout + row0 * C + h * 32;
#pragma unroll
    for (int t = 0; t < 25; ++t) {
      const int idx = lane + 32 * t;
      if (idx < 4 * BA_L * 4) {
        const int which = idx / (BA_L * 4), rem = idx - which * (BA_L * 4);
        const int r = rem >> 2, ch = rem & 3;
        const T* gp = which < 3 ? src + static_cast<long long>(r) * ld_qkv + which * C + ch * 8
                                : dsrc + static_cast<long long>(r) * C + ch * 8;
        ba_cp16(Qs + which * BA_MAT + ba_off(r, ch), gp);
      }
    }
    const int wy = w / g.nWx, wx = w - wy * g.nWx;
    const bool masked = g.shift > 0 && (wy == nWy - 1 || wx == g.nWx - 1);   // warp-uniform
    if (masked) {
      region_s[lane] = static_cast<int8_t>(win_region(g, w, lane));
      if (lane + 32 < BA_L) region_s[lane + 32] = static_cast<int8_t>(win_region(g, w, lane + 32));
    }
    ba_cp_wait();
    __syncwarp();

    // ------------------------------------------------ pass A: query-row tiles -> dQ, dbias, (L, D) per row
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int m0 = mt * 16;
      const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
      float s[7][4], dp[7][4];
      ba_tile_abt<T>(s, Qs, m0, Ks, lane);
      ba_tile_abt<T>(dp, Os, m0, Vs, lane);
      const int rr0 = r0 < BA_L ? r0 : BA_L - 1, rr1 = r1 < BA_L ? r1 : BA_L - 1;
      const int reg0 = masked ? region_s[rr0] : 0, reg1 = masked ? region_s[rr1] : 0;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = j * 8 + qc + e;
          float b0 = -INFINITY, b1 = -INFINITY;
          if (c < BA_L) {
            b0 = bias_s[rr0 * BA_L + c];
            b1 = bias_s[rr1 * BA_L + c];
            if (masked) {
              const int rc = region_s[c];
              if (rc != reg0) b0 += -100.0f;
              if (rc != reg1) b1 += -100.0f;
            }
          }
          s[j][e] = fmaf(s[j][e], scale, b0);
          s[j][2 + e] = fmaf(s[j][2 + e], scale, b1);
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
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[j][0] = __expf(s[j][0] - mx0); s[j][1] = __expf(s[j][1] - mx0);
        s[j][2] = __expf(s[j][2] - mx1); s[j][3] = __expf(s[j][3] - mx1);
        sum0 += s[j][0] + s[j][1];
        sum1 += s[j][2] + s[j][3];
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;          // P
        d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
        d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      if ((lane & 3) == 0) {
        rowL[r0] = mx0 + __logf(sum0); rowD[r0] = d0;
        rowL[r1] = mx1 + __logf(sum1); rowD[r1] = d1;
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[j][0] *= dp[j][0] - d0; s[j][1] *= dp[j][1] - d0;                          // dS
        s[j][2] *= dp[j][2] - d1; s[j][3] *= dp[j][3] - d1;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = j * 8 + qc + e;
          if (c < BA_L) {
            if (r0 < BA_L) atomicAdd(&dbias_s[r0 * BA_L + c], s[j][e]);
            if (r1 < BA_L) atomicAdd(&dbias_s[r1 * BA_L + c], s[j][2 + e]);
          }
        }
      }
      float o[4][4];
      ba_tile_wb<T>(o, s, Ks, lane);                                                  // dQ = scale dS K
      ba_store_tile<T>(dqkv, ld_qkv, row0, r0, r1, h * 32, o, scale, lane);
    }
    __syncwarp();   // rowL / rowD of all 64 rows are visible to the whole warp

    // ------------------------------------------------ pass B: key-row tiles -> dK, dV  (transposed tiles recomputed)
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int m0 = mt * 16;
      const int r0 = m0 + (lane >> 2), r1 = r0 + 8;                                   // key rows
      float s[7][4], dp[7][4];
      ba_tile_abt<T>(s, Ks, m0, Qs, lane);                                            // S^T (unscaled)
      ba_tile_abt<T>(dp, Vs, m0, Os, lane);                                           // dP^T
      const bool ok0 = r0 < BA_L, ok1 = r1 < BA_L;
      const int rr0 = ok0 ? r0 : BA_L - 1, rr1 = ok1 ? r1 : BA_L - 1;
      const int reg0 = masked ? region_s[rr0] : 0, reg1 = masked ? region_s[rr1] : 0;
      float pt[7][4];                                                                 // P^T, kept for dV
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int c0 = j * 8 + qc;                                                    // query columns c0, c0 + 1
        const float2 Lc = *reinterpret_cast<const float2*>(rowL + c0);
        const float2 Dc = *reinterpret_cast<const float2*>(rowD + c0);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = c0 + e;
          const float L = e ? Lc.y : Lc.x, D = e ? Dc.y : Dc.x;
          float b0 = -INFINITY, b1 = -INFINITY;
          if (c < BA_L) {
            if (ok0) b0 = bias_s[c * BA_L + rr0];
            if (ok1) b1 = bias_s[c * BA_L + rr1];
            if (masked) {
              const int rc = region_s[c];
              if (rc != reg0) b0 += -100.0f;
              if (rc != reg1) b1 += -100.0f;
            }
          }
          const float p0 = __expf(fmaf(s[j][e], scale, b0) - L);
          const float p1 = __expf(fmaf(s[j][2 + e], scale, b1) - L);
          pt[j][e] = p0; pt[j][2 + e] = p1;
          s[j][e] = p0 * (dp[j][e] - D);                                              // dS^T
          s[j][2 + e] = p1 * (dp[j][2 + e] - D);
        }
      }
      float o[4][4];
      ba_tile_wb<T>(o, s, Qs, lane);                                                  // dK = scale dS^T Q
      ba_store_tile<T>(dqkv, ld_qkv, row0, r0, r1, C + h * 32, o, scale, lane);
      ba_tile_wb<T>(o, pt, Os, lane);                                                 // dV = P^T dO
      ba_store_tile<T>(dqkv, ld_qkv, row0, r0, r1, 2 * C + h * 32, o, 1.0f, lane);
    }
    __syncwarp();   // all lanes are done with this window before the next cp.async overwrites it
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BA_L * BA_L; i += BA_WARPS * 32)
    atomicAdd(dbias + static_cast<long long>(h) * BA_L * BA_L + i, dbias_s[i]);
}

template <typename T>
static int launch_wab(const void* qkv, const void* dout, void* dqkv, const float* bias, float* dbias, int num_windows, int C,
                      int heads, const WinGeom& g, int nW, cudaStream_t stream) {
  static DeviceOnce once;
  auto kern = win_attn_bwd_kernel<T>;
  const int smem = 2 * BA_TABLE_BYTES + BA_WARPS * BA_WARP_BYTES;
  if (once.first()) {
    CSVIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  int per_head = (num_windows + BA_WARPS - 1) / BA_WARPS;
  const int cap = (148 * 3) / heads > 0 ? (148 * 3) / heads : 1;
  if (per_head > cap) per_head = cap;
  kern<<<per_head * heads, BA_WARPS * 32, smem, stream>>>(static_cast<const T*>(qkv), static_cast<const T*>(dout), static_cast<T*>(dqkv),
                                                          bias, dbias, num_windows, C, heads, g, nW, 0.17677669529663687f);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// Window-attention backward on packed tensors: qkv / dqkv [B*H*W, 3C], dout [B*H*W, C], window 7, head_dim 32, 16-bit dtype.
int launch_window_attention_bwd_mma(const void* qkv, const void* dout, void* dqkv, int dtype, const float* bias, float* dbias, int B,
                                    int H, int W, int C, int heads, int ws, int shift, cudaStream_t stream) {
  CSVIT_REQUIRE(ws == 7 && C == heads * 32, "window_attention_bwd(16-bit): window 7 and head_dim 32 only");
  CSVIT_REQUIRE(bias != nullptr && dbias != nullptr, "window_attention_bwd(16-bit): bias and dbias are required");
  const int nW = (H / ws) * (W / ws);
  if (B * nW <= 0) return 0;
  WinGeom g = make_geom(H, W, ws, shift);
  if (dtype == DT_BF16) return launch_wab<__nv_bfloat16>(qkv, dout, dqkv, bias, dbias, B * nW, C, heads, g, nW, stream);
  return launch_wab<__half>(qkv, dout, dqkv, bias, dbias, B * nW, C, heads, g, nW, stream);
}

}  // namespace csvit
