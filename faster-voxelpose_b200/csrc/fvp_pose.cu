// fvp_pose.cu - the tail of the JointLocalizationNet (joint_localization_net.py:15-62,84-98):
//   SoftArgmaxLayer (softmax(beta*x) over 4096 px, expectation of the plane coordinates, max prob),
//   WeightNet (weight_net.py:48-80: conv3x3(1->F)+BN, MaxPool2, ReLU, global average, MLP, sigmoid),
//   fuse_pose_preds, offset add, confidence mean and the final [B,P,J,5] packing
//   (faster_voxelpose.py:102-103).
// One CTA per (joint, person, plane) - 3*J*n CTAs, four per SM; the 64x64 map is staged once
// in shared memory and used by both the soft-argmax and the WeightNet; k_fuse blends the three planes per joint.  Expectations are accumulated in fp64 (joint
// coordinates are ~1e3 mm where one fp32 ulp is 1.2e-4 mm; see DESIGN.md "numerics").
#include "fvp_kernels.h"

namespace {

constexpr int PT = 256;            // threads
constexpr int MS = 68;             // smem row stride of the zero-bordered 66x66 map

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < PT / 32; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
#pragma unroll
  for (int w = 1; w < PT / 32; ++w) r += red[w];
  __syncthreads();
  return r;
}

template <int F>
__global__ void __launch_bounds__(PT, 4) k_pose_head(FvpGeom g, FvpPoseW w, const float* __restrict__ feat,
                                                  const FvpPerson* __restrict__ people,
                                                  const float* __restrict__ offsets, int n, float beta,
                                                  float* __restrict__ pose, float* __restrict__ maxw,
                                                  float* __restrict__ weights) {
  __shared__ float s_map[66 * MS];
  __shared__ float s_cw[F * 9], s_cb[F];
  __shared__ float s_red[PT / 32];
  __shared__ double s_dred[PT / 32];
  __shared__ float s_gap[PT / 32][F];
  __shared__ float s_feat[F], s_hid[128];

  const int j = blockIdx.x, person = blockIdx.y, q = blockIdx.z, tid = threadIdx.x;
  const int J = g.J;
  float off[3];
  if (people) {
    const FvpPerson pd = people[person];
    if (!pd.valid) return;
    off[0] = pd.offset[0]; off[1] = pd.offset[1]; off[2] = pd.offset[2];
  } else {
    off[0] = offsets[person * 3]; off[1] = offsets[person * 3 + 1]; off[2] = offsets[person * 3 + 2];
  }
  for (int i = tid; i < F * 9; i += PT) s_cw[i] = w.conv_w[i];
  for (int i = tid; i < F; i += PT) s_cb[i] = w.conv_b[i];
  for (int i = tid; i < 66 * MS; i += PT) s_map[i] = 0.f;
  __syncthreads();

  {
    const float* m = feat + (((size_t)q * n + person) * J + j) * 4096;
    // ---- stage the map; soft-argmax statistics -------------------------------------------------
    float t[16];
    float tmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = tid + PT * k;
      const float x = __ldg(m + i);
      s_map[((i >> 6) + 1) * MS + (i & 63) + 1] = x;
      t[k] = __fmul_rn(beta, x);
      tmax = fmaxf(tmax, t[k]);
    }
    const float gmax = block_max(tmax, s_red);    // also orders the s_map stores
    const float* ax0 = g.ind_axes + (q == 2 ? 64 : 0);      // first plane coordinate: x, x, y
    const float* ax1 = g.ind_axes + (q == 0 ? 64 : 128);    // second: y, z, z
    double se = 0.0, s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = tid + PT * k;
      const double e = (double)expf(__fsub_rn(t[k], gmax));
      se += e;
      s0 += e * (double)__ldg(ax0 + (i >> 6));
      s1 += e * (double)__ldg(ax1 + (i & 63));
    }
    se = block_sum(se, s_dred);
    s0 = block_sum(s0, s_dred);
    s1 = block_sum(s1, s_dred);

    // ---- WeightNet on the staged map ------------------------------------------------------------
    // each thread owns 4 pooled pixels (a 4x4 input patch each), two per pass: 32 patch registers instead of 64 keep the
    // kernel at 64 registers = 4 CTAs per SM, so the 3 * J * P = 450 CTAs of a batch-1 frame are ONE wave (592 slots; at 3 per
    // SM, 444 slots, six CTAs ran alone in a second wave).  Channels are a runtime loop so the patches stay in registers and
    // each channel's pooled sum is warp-reduced immediately.
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float in[2][4][4];
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int pix = tid + PT * (2 * pass + pp);              // pooled pixel 0..1023
        const int py = pix >> 5, px = pix & 31;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) in[pp][r][c] = s_map[(2 * py + r) * MS + 2 * px + c];
      }
#pragma unroll 2
      for (int c = 0; c < F; ++c) {
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = s_cw[c * 9 + i];
        const float cb = s_cb[c];
        float v = 0.f;
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          float best = -INFINITY;
#pragma unroll
          for (int oy = 0; oy < 2; ++oy)
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
              float a = cb;
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) a = fmaf(k[r * 3 + cc], in[pp][oy + r][ox + cc], a);
              best = fmaxf(best, a);
            }
          v += fmaxf(best, 0.f);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_gap[tid >> 5][c] = pass ? s_gap[tid >> 5][c] + v : v;
      }
    }
    __syncthreads();
    if (tid < F) {
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < PT / 32; ++wv) v += s_gap[wv][tid];
      s_feat[tid] = v * (1.0f / 1024.0f);
    }
    __syncthreads();
    if (tid < w.hidden) {
      float a = w.fc1_b[tid];
      for (int c = 0; c < F; ++c) a = fmaf(w.fc1_w[tid * F + c], s_feat[c], a);
      s_hid[tid] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (tid == 0) {
      float a = w.fc2_b[0];
      for (int c = 0; c < w.hidden; ++c) a = fmaf(w.fc2_w[c], s_hid[c], a);
      const float sg = 1.0f / (1.0f + expf(-a));
      const float p0 = (float)(s0 / se), p1 = (float)(s1 / se);
      const int o0 = q == 2 ? 1 : 0, o1 = q == 0 ? 1 : 2;   // offsets: (x,y) (x,z) (y,z)
      const size_t o = ((size_t)q * n + person) * J + j;
      maxw[o] = (float)(1.0 / se);
      weights[o] = sg;
      pose[o * 2] = __fadd_rn(p0, off[o0]);
      pose[o * 2 + 1] = __fadd_rn(p1, off[o1]);
    }
  }
}

// fuse_pose_preds (joint_localization_net.py:50-59): one thread per (person, joint) normalises the pair of plane
// weights of every axis and blends the two plane coordinates
__global__ void k_fuse(const FvpPerson* __restrict__ people, int n, int J, const float* __restrict__ pose,
                       const float* __restrict__ weights, float* __restrict__ fused) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // person * J + joint
  if (i >= n * J) return;
  if (people && !people[i / J].valid) return;
  const size_t nj = (size_t)n * J;
  const float wxy = weights[i], wxz = weights[nj + i], wyz = weights[2 * nj + i];
  const float* pxy = pose + (size_t)i * 2;
  const float* pxz = pose + (nj + i) * 2;
  const float* pyz = pose + (2 * nj + i) * 2;
  const float sx = __fadd_rn(wxy, wxz), sy = __fadd_rn(wxy, wyz), sz = __fadd_rn(wxz, wyz);
  const float x = __fadd_rn(__fmul_rn(__fdiv_rn(wxy, sx), pxy[0]), __fmul_rn(__fdiv_rn(wxz, sx), pxz[0]));
  const float y = __fadd_rn(__fmul_rn(__fdiv_rn(wxy, sy), pxy[1]), __fmul_rn(__fdiv_rn(wyz, sy), pyz[0]));
  const float z = __fadd_rn(__fmul_rn(__fdiv_rn(wxz, sz), pxz[1]), __fmul_rn(__fdiv_rn(wyz, sz), pyz[1]));
  float* f = fused + (size_t)i * 3;
  f[0] = x; f[1] = y; f[2] = z;
}

// one thread per (slot, joint): confidence + final packing
__global__ void k_finalize(FvpGeom g, const FvpPerson* __restrict__ people, const float* __restrict__ maxw,
                           const float* __restrict__ pose, const float* __restrict__ fused,
                           float* __restrict__ centers, int n, float* __restrict__ conf_out,
                           float* __restrict__ fused_poses, float* __restrict__ plane_poses,
                           float* __restrict__ centers_out) {
  const int slot = blockIdx.x, j = threadIdx.x, J = g.J;
  __shared__ float s_conf;
  const bool valid = people[slot].valid != 0;
  if (j == 0) {
    float c = centers[(size_t)slot * 7 + 4];
    if (valid) {                                  // mean over (plane, joint) of the max soft-max weight
      float s = 0.f;
      for (int q = 0; q < 3; ++q)
        for (int jj = 0; jj < J; ++jj) s += maxw[((size_t)q * n + slot) * J + jj];
      c = s / (float)(3 * J);
      centers[(size_t)slot * 7 + 4] = c;          // joint_localization_net.py:98 (write through the alias)
    }
    s_conf = c;
    if (conf_out) conf_out[slot] = valid ? c : 0.f;
  }
  __syncthreads();
  if (j < 7 && centers_out) centers_out[(size_t)slot * 7 + j] = (j == 4) ? s_conf : centers[(size_t)slot * 7 + j];
  if (j >= J) return;
  if (fused_poses) {
    float* o = fused_poses + ((size_t)slot * J + j) * 5;
    const float* f = fused + ((size_t)slot * J + j) * 3;
    o[0] = valid ? f[0] : 0.f;
    o[1] = valid ? f[1] : 0.f;
    o[2] = valid ? f[2] : 0.f;
    o[3] = centers[(size_t)slot * 7 + 3];
    o[4] = s_conf;
  }
  if (plane_poses) {
    for (int q = 0; q < 3; ++q) {
      const size_t o = (((size_t)q * n + slot) * J + j) * 2;
      plane_poses[o] = valid ? pose[o] : 0.f;
      plane_poses[o + 1] = valid ? pose[o + 1] : 0.f;
    }
  }
}

}  // namespace

void fvp_launch_pose_head(const FvpGeom& g, const FvpPoseW& w, const float* d_feat, const FvpPerson* d_people,
                          const float* d_offset, int n, float beta, float* d_pose, float* d_maxw, float* d_weights,
                          float* d_fused, cudaStream_t st) {
  dim3 grid(g.J, n, 3);
  k_pose_head<32><<<grid, PT, 0, st>>>(g, w, d_feat, d_people, d_offset, n, beta, d_pose, d_maxw, d_weights);
  k_fuse<<<fvp_cdiv(n * g.J, 128), 128, 0, st>>>(d_people, n, g.J, d_pose, d_weights, d_fused);
}

void fvp_launch_finalize(const FvpGeom& g, const FvpPerson* d_people, const float* d_maxw, const float* d_pose,
                         const float* d_fused, float* d_centers, int batch, float* d_conf, float* d_fused_poses,
                         float* d_plane_poses, float* d_centers_out, cudaStream_t st) {
  const int n = batch * g.P;
  k_finalize<<<n, 32, 0, st>>>(g, d_people, d_maxw, d_pose, d_fused, d_centers, n, d_conf, d_fused_poses,
                               d_plane_poses, d_centers_out);
}
