// The fp32 tail of Poser.predict_batch after the three output heads, as two kernels instead of ~180 elementwise / small-GEMM
// launches (7 % of the Swin-B batch-256 step in round 1):
//
//   rot6d_to_axis_angle   pose_6d [n, 6] -> axis-angle [n, 3]: Gram-Schmidt 6D -> rotation matrix -> quaternion (best-conditioned
//                         candidate, real part >= 0) -> axis-angle, with the branch structure of ref:cs_vit/utils/geometry.py:111-132,
//                         150-223, 258-298 (called at ref:cs_vit/net/ti_poser.py:529-534) because it is observable near pi.
//   mano_fk               everything of Poser._pose_fk (ref:cs_vit/net/ti_poser.py:561-607) for one sample per CTA: linear-blend
//                         skinning of the MANO layer (shape blend -> joint regression -> Rodrigues -> 16-joint kinematic chain ->
//                         optional pose-corrective blend -> skinning), the 21-joint regression J_regressor_mano, the mean bone
//                         length, de-normalisation of the root translation to millimetres and the root-relative camera-space
//                         joints / vertices.  The LBS follows the layer the model was given: the seeded stand-in of this repo
//                         (cs_vit/utils/mano_standin.py, rodrigues_mode 0) or the published smplx `lbs` (rodrigues_mode 1, with
//                         posedirs and the hand-pose mean) - the latter is UNPINNED here: smplx and the MANO files are not in this image.
// Everything is fp32; vertices stay in shared memory between the skinning and the final write.
#include "common.cuh"
#include "errors.h"
#include "rowops.cuh"

namespace csvit {

// ---------------------------------------------------------------------------------------------- 6D -> axis-angle
__global__ void __launch_bounds__(256)
rot6d_to_axis_angle_kernel(const float* __restrict__ d6, float* __restrict__ aa, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = d6 + 6 * i;
  const float a1x = p[0], a1y = p[1], a1z = p[2], a2x = p[3], a2y = p[4], a2z = p[5];
  // F.normalize: x / max(|x|, 1e-12)
  float inv = 1.0f / fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  const float b1x = a1x * inv, b1y = a1y * inv, b1z = a1z * inv;
  const float dot = b1x * a2x + b1y * a2y + b1z * a2z;
  float ux = a2x - dot * b1x, uy = a2y - dot * b1y, uz = a2z - dot * b1z;
  inv = 1.0f / fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);
  const float b2x = ux * inv, b2y = uy * inv, b2z = uz * inv;
  const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  // rows of the matrix are (b1, b2, b3): m[r][c]
  const float m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z, m20 = b3x, m21 = b3y, m22 = b3z;
  float sq[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
  float qa[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) qa[k] = sq[k] > 0.0f ? sqrtf(sq[k]) : 0.0f;
  const float a = m21 - m12, b = m02 - m20, c = m10 - m01, e = m10 + m01, f = m02 + m20, g = m12 + m21;
  int best = 0;                       // torch.argmax: first index of the maximum
#pragma unroll
  for (int k = 1; k < 4; ++k) if (qa[k] > qa[best]) best = k;
  float q0, q1, q2, q3;
  const float den = 2.0f * fmaxf(qa[best], 0.1f);
  if (best == 0) { q0 = qa[0] * qa[0]; q1 = a; q2 = b; q3 = c; }
  else if (best == 1) { q0 = a; q1 = qa[1] * qa[1]; q2 = e; q3 = f; }
  else if (best == 2) { q0 = b; q1 = e; q2 = qa[2] * qa[2]; q3 = g; }
  else { q0 = c; q1 = f; q2 = g; q3 = qa[3] * qa[3]; }
  q0 /= den; q1 /= den; q2 /= den; q3 /= den;
  if (q0 < 0.0f) { q0 = -q0; q1 = -q1; q2 = -q2; q3 = -q3; }
  const float nrm = sqrtf(q1 * q1 + q2 * q2 + q3 * q3);
  const float half = atan2f(nrm, q0);
  // q[1:] / (0.5 * sinc(half / pi)), sinc(x) = sin(pi x) / (pi x), sinc(0) = 1
  const float s = half == 0.0f ? 0.5f : 0.5f * sinf(half) / half;
  float* o = aa + 3 * i;
  o[0] = q1 / s; o[1] = q2 / s; o[2] = q3 / s;
}

int launch_rot6d_to_axis_angle(const float* d6, float* aa, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  rot6d_to_axis_angle_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(d6, aa, n);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- MANO FK + de-normalisation
constexpr int MK_THREADS = 256;
constexpr int MK_V = 778, MK_J = 16, MK_JO = 21, MK_EDGES = 20;

struct ManoParams {
  const float* pose;         // [n, 48] axis-angle: global orientation + 15 hand joints
  const float* betas;        // [n, 10]
  const float* root_norm;    // [n, 3] predicted root translation, in units of the mean bone length
  const float* v_template;   // [778, 3]
  const float* shapedirs;    // [778, 3, 10]
  const float* posedirs;     // [135, 778 * 3] or nullptr
  const float* pose_mean;    // [45] added to the hand pose, or nullptr
  const float* j_regressor;  // [16, 778]
  const float* lbs_weights;  // [778, 16]
  const float* j_out;        // [21, 778]  J_regressor_mano of the model
  float* joint_cam;          // [n, 21, 3]
  float* verts_cam;          // [n, 778, 3]
  float* root_transl;        // [n, 3]
  int n;
  int rodrigues_mode;        // 0: theta = sqrt(|a|^2 + 1e-16) (stand-in); 1: theta = |a + 1e-8| (smplx batch_rodrigues)
  int parents[MK_J];
  int edge_a[MK_EDGES], edge_b[MK_EDGES];
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {      // all threads get the sum; scratch: [MK_THREADS / 32]
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < MK_THREADS / 32; ++w) t += scratch[w];
  return t;
}

__global__ void __launch_bounds__(MK_THREADS)
mano_fk_kernel(ManoParams p) {
  __shared__ float vs[MK_V * 3];            // v_shaped, then v_posed, finally the skinned vertices
  __shared__ float rot[MK_J][9], wr[MK_J][9], wt[MK_J][3], rest[MK_J][3], jn[MK_J][3];
  __shared__ float jo[MK_JO][3];
  __shared__ float beta[10], scratch[MK_THREADS / 32], red[MK_JO * 3];
  const int s = blockIdx.x, tid = threadIdx.x;
  if (tid < 10) beta[tid] = p.betas[static_cast<long long>(s) * 10 + tid];
  if (tid < MK_J) {      // Rodrigues per joint
    float ax = p.pose[static_cast<long long>(s) * 48 + 3 * tid], ay = p.pose[static_cast<long long>(s) * 48 + 3 * tid + 1],
          az = p.pose[static_cast<long long>(s) * 48 + 3 * tid + 2];
    if (p.pose_mean != nullptr && tid > 0) { ax += p.pose_mean[3 * (tid - 1)]; ay += p.pose_mean[3 * (tid - 1) + 1]; az += p.pose_mean[3 * (tid - 1) + 2]; }
    float theta;
    if (p.rodrigues_mode == 0) theta = sqrtf(ax * ax + ay * ay + az * az + 1e-16f);
    else { const float bx = ax + 1e-8f, by = ay + 1e-8f, bz = az + 1e-8f; theta = sqrtf(bx * bx + by * by + bz * bz); }
    const float kx = ax / theta, ky = ay / theta, kz = az / theta;
    const float sn = sinf(theta), cs = 1.0f - cosf(theta);
    // I + sin K + (1 - cos) K^2,  K = [[0,-kz,ky],[kz,0,-kx],[-ky,kx,0]]
    rot[tid][0] = 1.0f + cs * (-kz * kz - ky * ky); rot[tid][1] = -sn * kz + cs * (kx * ky);         rot[tid][2] = sn * ky + cs * (kx * kz);
    rot[tid][3] = sn * kz + cs * (kx * ky);         rot[tid][4] = 1.0f + cs * (-kz * kz - kx * kx); rot[tid][5] = -sn * kx + cs * (ky * kz);
    rot[tid][6] = -sn * ky + cs * (kx * kz);        rot[tid][7] = sn * kx + cs * (ky * kz);         rot[tid][8] = 1.0f + cs * (-ky * ky - kx * kx);
  }
  __syncthreads();
  // shape blend
  for (int i = tid; i < MK_V * 3; i += MK_THREADS) {
    float v = p.v_template[i];
#pragma unroll
    for (int k = 0; k < 10; ++k) v = fmaf(p.shapedirs[i * 10 + k], beta[k], v);
    vs[i] = v;
  }
  __syncthreads();
  // rest joints: J_regressor [16, 778] x v_shaped
  for (int j = 0; j < MK_J; ++j) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int v = tid; v < MK_V; v += MK_THREADS) {
      const float w = p.j_regressor[j * MK_V + v];
      a0 = fmaf(w, vs[3 * v], a0); a1 = fmaf(w, vs[3 * v + 1], a1); a2 = fmaf(w, vs[3 * v + 2], a2);
    }
    a0 = block_sum(a0, scratch); a1 = block_sum(a1, scratch); a2 = block_sum(a2, scratch);
    if (tid == 0) { jn[j][0] = a0; jn[j][1] = a1; jn[j][2] = a2; }
  }
  __syncthreads();
  // pose-corrective blend (smplx): v_posed = v_shaped + (R[1:] - I).flatten() @ posedirs
  if (p.posedirs != nullptr) {
    for (int i = tid; i < MK_V * 3; i += MK_THREADS) {
      float v = vs[i];
      for (int f = 0; f < 135; ++f) {
        const int j = 1 + f / 9, e = f % 9;
        const float pf = rot[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        v = fmaf(pf, p.posedirs[static_cast<long long>(f) * (MK_V * 3) + i], v);
      }
      vs[i] = v;      // (each thread touches only its own entries: no barrier needed before the write)
    }
  }
  // kinematic chain (one thread: 16 dependent 3x3 products)
  if (tid == 0) {
    for (int e = 0; e < 9; ++e) wr[0][e] = rot[0][e];
    for (int e = 0; e < 3; ++e) wt[0][e] = jn[0][e];
    for (int j = 1; j < MK_J; ++j) {
      const int par = p.parents[j];
      const float rx = jn[j][0] - jn[par][0], ry = jn[j][1] - jn[par][1], rz = jn[j][2] - jn[par][2];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
          wr[j][3 * r + c] = wr[par][3 * r] * rot[j][c] + wr[par][3 * r + 1] * rot[j][3 + c] + wr[par][3 * r + 2] * rot[j][6 + c];
        wt[j][r] = wt[par][r] + wr[par][3 * r] * rx + wr[par][3 * r + 1] * ry + wr[par][3 * r + 2] * rz;
      }
    }
    for (int j = 0; j < MK_J; ++j)
      for (int r = 0; r < 3; ++r)
        rest[j][r] = wt[j][r] - (wr[j][3 * r] * jn[j][0] + wr[j][3 * r + 1] * jn[j][1] + wr[j][3 * r + 2] * jn[j][2]);
  }
  __syncthreads();
  // skinning, in place
  for (int v = tid; v < MK_V; v += MK_THREADS) {
    float R[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, t[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < MK_J; ++j) {
      const float w = p.lbs_weights[v * MK_J + j];
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = fmaf(w, wr[j][e], R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) t[e] = fmaf(w, rest[j][e], t[e]);
    }
    const float x = vs[3 * v], y = vs[3 * v + 1], z = vs[3 * v + 2];
    vs[3 * v] = R[0] * x + R[1] * y + R[2] * z + t[0];
    vs[3 * v + 1] = R[3] * x + R[4] * y + R[5] * z + t[1];
    vs[3 * v + 2] = R[6] * x + R[7] * y + R[8] * z + t[2];
  }
  __syncthreads();
  // 21 output joints: J_regressor_mano [21, 778] x vertices   (ref:cs_vit/net/ti_poser.py:582)
  for (int j = 0; j < MK_JO; ++j) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int v = tid; v < MK_V; v += MK_THREADS) {
      const float w = p.j_out[j * MK_V + v];
      a0 = fmaf(w, vs[3 * v], a0); a1 = fmaf(w, vs[3 * v + 1], a1); a2 = fmaf(w, vs[3 * v + 2], a2);
    }
    a0 = block_sum(a0, scratch); a1 = block_sum(a1, scratch); a2 = block_sum(a2, scratch);
    if (tid == 0) { jo[j][0] = a0; jo[j][1] = a1; jo[j][2] = a2; }
  }
  __syncthreads();
  // mean bone length (metres) -> root translation in mm; root-relative joints / vertices in mm + root   (ref :584-604)
  if (tid == 0) {
    float len = 0.f;
    for (int e = 0; e < MK_EDGES; ++e) {
      const float dx = jo[p.edge_a[e]][0] - jo[p.edge_b[e]][0], dy = jo[p.edge_a[e]][1] - jo[p.edge_b[e]][1],
                  dz = jo[p.edge_a[e]][2] - jo[p.edge_b[e]][2];
      len += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    const float mean_len = 1e3f * (len / float(MK_EDGES));
    for (int e = 0; e < 3; ++e) {
      const float r = p.root_norm[static_cast<long long>(s) * 3 + e] * mean_len;
      red[e] = r;
      p.root_transl[static_cast<long long>(s) * 3 + e] = r;
    }
  }
  __syncthreads();
  const float r0 = red[0], r1 = red[1], r2 = red[2];
  const float o0 = jo[0][0], o1 = jo[0][1], o2 = jo[0][2];
  if (tid < MK_JO) {
    float* o = p.joint_cam + (static_cast<long long>(s) * MK_JO + tid) * 3;
    o[0] = (jo[tid][0] - o0) * 1e3f + r0; o[1] = (jo[tid][1] - o1) * 1e3f + r1; o[2] = (jo[tid][2] - o2) * 1e3f + r2;
  }
  float* vo = p.verts_cam + static_cast<long long>(s) * MK_V * 3;
  for (int i = tid; i < MK_V * 3; i += MK_THREADS) {
    const int c = i % 3;
    vo[i] = (vs[i] - (c == 0 ? o0 : (c == 1 ? o1 : o2))) * 1e3f + (c == 0 ? r0 : (c == 1 ? r1 : r2));
  }
}

int launch_mano_fk(const float* pose, const float* betas, const float* root_norm, const float* v_template, const float* shapedirs,
                   const float* posedirs, const float* pose_mean, const float* j_regressor, const float* lbs_weights, const float* j_out,
                   const int* parents16, const int* edges40, int rodrigues_mode, float* joint_cam, float* verts_cam, float* root_transl,
                   int n, cudaStream_t stream) {
  if (n <= 0) return 0;
  ManoParams p{};
  p.pose = pose; p.betas = betas; p.root_norm = root_norm; p.v_template = v_template; p.shapedirs = shapedirs; p.posedirs = posedirs;
  p.pose_mean = pose_mean; p.j_regressor = j_regressor; p.lbs_weights = lbs_weights; p.j_out = j_out;
  p.joint_cam = joint_cam; p.verts_cam = verts_cam; p.root_transl = root_transl; p.n = n; p.rodrigues_mode = rodrigues_mode;
  CSVIT_REQUIRE(parents16[0] < 0, "mano_fk: joint 0 must be the root (parent -1)");
  for (int j = 0; j < MK_J; ++j) {
    CSVIT_REQUIRE(j == 0 || (parents16[j] >= 0 && parents16[j] < j), "mano_fk: parent of joint %d must precede it (got %d)", j, parents16[j]);
    p.parents[j] = parents16[j];
  }
  for (int e = 0; e < MK_EDGES; ++e) {
    CSVIT_REQUIRE(edges40[2 * e] >= 0 && edges40[2 * e] < MK_JO && edges40[2 * e + 1] >= 0 && edges40[2 * e + 1] < MK_JO, "mano_fk: bad skeleton edge %d", e);
    p.edge_a[e] = edges40[2 * e]; p.edge_b[e] = edges40[2 * e + 1];
  }
  mano_fk_kernel<<<n, MK_THREADS, 0, stream>>>(p);
  CSVIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace csvit
