// On-device crop-and-resize of the hand region: full frames already in HBM -> the [N,3,S,S] fp32
// patches Poser.predict_batch takes, replacing the data sets' per-sample CPU crop
//     crop_tensor_with_square_box(img, bbox_tight, expansion_ratio, S)        ref:cs_vit/utils/img.py:339-390
// (tight xyxy box -> square box on the longer side, scaled about the centre by expansion_ratio, ref :358-370; then
// kornia crop_and_resize(mode='bilinear', padding_mode='zeros', align_corners=True), ref :372-388, which for an axis-aligned box
// samples output pixel (u, v) at source (x1 + u (x2-x1)/(S-1), y1 + v (y2-y1)/(S-1)) with zeros outside the frame).
// One thread per output pixel computes the three channels: four gathers per channel from the frame (L2-resident neighbourhood),
// one coalesced fp32 store per channel.  HBM-bound: 12 S^2 bytes written per crop + the touched part of the frame read once.
// Frames: fp32 CHW in [0,1] (what the reference's loaders hand to the crop) or uint8 HWC as decoded (scaled by 1/255 here:
// bilinear interpolation is linear, so interpolating the bytes and scaling is the same as the reference's scale-then-interpolate).
#include "common.cuh"
#include "errors.h"
#include "rowops.cuh"

namespace csvit {

struct CropParams {
  const void* frames;
  const float* boxes;      // [N,4] xyxy: tight boxes (expansion > 0) or the final crop boxes (expansion <= 0)
  float* square_out;       // [N,4] the boxes actually cropped, or nullptr
  float* out;              // [N,3,S,S]
  int N, H, W, S;
  float expansion;
};

template <bool U8>
__device__ __forceinline__ float crop_fetch(const CropParams& p, int n, int c, int y, int x) {
  if (x < 0 || y < 0 || x >= p.W || y >= p.H) return 0.0f;          // padding_mode='zeros'
  if (U8) return float(static_cast<const uint8_t*>(p.frames)[((static_cast<long long>(n) * p.H + y) * p.W + x) * 3 + c]) * (1.0f / 255.0f);
  return static_cast<const float*>(p.frames)[((static_cast<long long>(n) * 3 + c) * p.H + y) * p.W + x];
}

template <bool U8>
__global__ void __launch_bounds__(256)
crop_resize_kernel(CropParams p) {
  const long long total = static_cast<long long>(p.N) * p.S * p.S;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int u = static_cast<int>(i % p.S);
    const long long t = i / p.S;
    const int v = static_cast<int>(t % p.S), n = static_cast<int>(t / p.S);
    float x1 = p.boxes[4 * n], y1 = p.boxes[4 * n + 1], x2 = p.boxes[4 * n + 2], y2 = p.boxes[4 * n + 3];
    if (p.expansion > 0.0f) {        // ref:cs_vit/utils/img.py:358-370
      const float cx = (x1 + x2) / 2, cy = (y1 + y2) / 2;
      const float half = fmaxf(x2 - x1, y2 - y1) * p.expansion / 2;
      x1 = cx - half; x2 = cx + half; y1 = cy - half; y2 = cy + half;
    }
    if (p.square_out != nullptr && u == 0 && v == 0) {
      p.square_out[4 * n] = x1; p.square_out[4 * n + 1] = y1; p.square_out[4 * n + 2] = x2; p.square_out[4 * n + 3] = y2;
    }
    const float den = p.S > 1 ? float(p.S - 1) : 1.0f;
    const float sx = x1 + float(u) * (x2 - x1) / den, sy = y1 + float(v) * (y2 - y1) / den;
    const float fx = floorf(sx), fy = floorf(sy);
    const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
    const float ax = sx - fx, ay = sy - fy;
    const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay), w10 = (1.0f - ax) * ay, w11 = ax * ay;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float val = w00 * crop_fetch<U8>(p, n, c, iy, ix) + w01 * crop_fetch<U8>(p, n, c, iy, ix + 1) +
                        w10 * crop_fetch<U8>(p, n, c, iy + 1, ix) + w11 * crop_fetch<U8>(p, n, c, iy + 1, ix + 1);
      p.out[((static_cast<long long>(n) * 3 + c) * p.S + v) * p.S + u] = val;
    }
  }
}

int launch_crop_resize(const void* frames, int frames_u8, int N, int H, int W, const float* boxes, float expansion, float* square_out,
                       float* out, int S, cudaStream_t stream) {
  CSVIT_REQUIRE(N >= 0 && H > 0 && W > 0 && S > 0, "crop_resize: bad shape N=%d H=%d W=%d S=%d", N, H, W, S);
  if (N == 0) return 0;
  CropParams p{frames, boxes, square_out, out, N, H, W, S, expansion};
  const long long total = static_cast<long long>(N) * S * S;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (frames_u8) crop_resize_kernel<true><<<static_cast<int>(blocks), 256, 0, stream>>>(p);
  else crop_resize_kernel<false><<<static_cast<int>(blocks), 256, 0, stream>>>(p);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace csvit
