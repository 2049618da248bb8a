// fvp_render.cu - N1: the heat-map renderer that feeds the hot path when the reference runs with
// TEST_HEATMAP_SRC 'pred' / 'gt': JointsDataset.generate_input_heatmap (lib/dataset/JointsDataset.py:271-337) and
// compute_human_scale (:197-203), evaluation branch (no augmentation).
//
// The reference loops persons x joints on the CPU and pastes a (6*sigma+1)^2 float64 Gaussian patch per joint into a
// float32 map with np.maximum.  Here:
//   k_hm_prepare : one thread per (frame, view, person): person scale -> sigma -> per-joint patch descriptor
//                  (integer patch origin with the reference's int() truncation toward zero, clipped window, centre index,
//                  denominator), all in IEEE float64 exactly as NumPy evaluates them;
//   k_hm_render  : four consecutive pixels of a (frame, view, joint) plane per thread (one 16-byte store), a CTA = 1024
//                  pixels; one warp first keeps the descriptors whose window meets the CTA's rows (ballot compaction),
//                  then every pixel takes the maximum over them
//                  of float(exp(-((gx-c0)^2 + (gy-c0)^2) / den)) - float64 argument and exp, rounded once to float32,
//                  which is what the reference's float32 assignment does.  The map is written exactly once (zeros
//                  included): the kernel is bound by the 4*V*J*H*W bytes it stores.
#include "fvp_kernels.h"

namespace {

struct HmPatch {        // one (person, joint) Gaussian patch of one view
  int ulx, uly;         // patch origin in the map (may be negative)
  int x0, x1, y0, y1;   // window of the map it touches; x0 >= x1 = nothing to draw
  double c0;            // centre index inside the patch: size // 2
  double den;           // 2 * cur_sigma^2
};

// trunc toward zero like Python's int(float)
__device__ __forceinline__ int py_int(double v) { return (int)v; }

__global__ void k_hm_prepare(const double* __restrict__ joints, const int* __restrict__ num, const unsigned char* __restrict__ vis,
                             int total_views, int max_people, int J, int W, int H, double stride_x, double stride_y,
                             double sigma, HmPatch* __restrict__ patches) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (frame*V + view) * max_people + person
  if (i >= total_views * max_people) return;
  const int bv = i / max_people, n = i - bv * max_people;
  HmPatch* out = patches + ((size_t)bv * J) * max_people + n;            // [bv][j][n]
  HmPatch none;
  none.ulx = none.uly = none.x0 = none.x1 = none.y0 = none.y1 = 0;
  none.c0 = none.den = 1.0;
  if (n >= num[bv]) {
    for (int j = 0; j < J; ++j) out[(size_t)j * max_people] = none;
    return;
  }
  const double* p = joints + (size_t)i * J * 2;
  // compute_human_scale on joints / feat_stride with every joint counted (JointsDataset.py:278-279)
  double minx = p[0] / stride_x, maxx = minx, miny = p[1] / stride_y, maxy = miny;
  for (int j = 1; j < J; ++j) {
    const double x = p[2 * j] / stride_x, y = p[2 * j + 1] / stride_y;
    minx = fmin(minx, x); maxx = fmax(maxx, x);
    miny = fmin(miny, y); maxy = fmax(maxy, y);
  }
  const double ext = fmax(maxy - miny, maxx - minx);
  double hs = ext * ext;                                                 // (...)**2
  hs = fmin(fmax(hs, 1.0 / 4 * 96 * 96), 4.0 * 96 * 96);                 // np.clip
  hs = 2.0 * hs;
  const double cur_sigma = sigma * sqrt(hs / (96.0 * 96.0));
  const double tmp = cur_sigma * 3.0;
  const double size = 2.0 * tmp + 1.0;
  const double c0 = floor(size / 2.0);                                   // size // 2
  const double den = 2.0 * (cur_sigma * cur_sigma);
  for (int j = 0; j < J; ++j) {
    HmPatch d = none;
    const bool seen = vis == nullptr || vis[(size_t)i * J + j] != 0;
    if (seen) {
      const int mu_x = py_int(p[2 * j] / stride_x), mu_y = py_int(p[2 * j + 1] / stride_y);
      const int ulx = py_int((double)mu_x - tmp), uly = py_int((double)mu_y - tmp);
      const int brx = py_int((double)mu_x + tmp + 1.0), bry = py_int((double)mu_y + tmp + 1.0);
      if (!(ulx >= W || uly >= H || brx < 0 || bry < 0)) {
        d.ulx = ulx; d.uly = uly;
        d.x0 = max(0, ulx); d.x1 = min(brx, W);
        d.y0 = max(0, uly); d.y1 = min(bry, H);
        d.c0 = c0; d.den = den;
        if (d.y0 >= d.y1) d.x1 = d.x0;                                   // empty window
      }
    }
    out[(size_t)j * max_people] = d;
  }
}

constexpr int HM_MAXP = FVP_MAX_PEOPLE;
static_assert(HM_MAXP <= 32, "the descriptor filter is one warp wide");

// VEC pixels per thread (4 when W % 4 == 0: one 16-byte store).  A CTA covers 256*VEC consecutive pixels of one plane.
template <int VEC>
__global__ void __launch_bounds__(256) k_hm_render(const HmPatch* __restrict__ patches, int max_people, int J, int W, int H,
                                                    float* __restrict__ out) {
  __shared__ HmPatch s_p[HM_MAXP];
  __shared__ int s_n;
  const int plane = blockIdx.x;                               // (frame*V + view) * J + joint
  const int i0 = blockIdx.y * 256 * VEC, i = i0 + threadIdx.x * VEC;
  if (threadIdx.x < 32) {                                     // keep the patches whose window meets this CTA's rows
    const int row_lo = i0 / W, row_hi = min(H - 1, (i0 + 256 * VEC - 1) / W);
    HmPatch d;
    bool keep = false;
    if ((int)threadIdx.x < max_people) {
      d = patches[(size_t)plane * max_people + threadIdx.x];
      keep = d.x0 < d.x1 && d.y0 <= row_hi && d.y1 > row_lo;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) s_p[__popc(m & ((1u << threadIdx.x) - 1u))] = d;
    if (threadIdx.x == 0) s_n = __popc(m);
  }
  __syncthreads();
  if (i >= W * H) return;
  const int py = i / W, px0 = i - py * W;                     // VEC > 1: W % VEC == 0, so the pixels share the row
  float v[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) v[e] = 0.0f;
  const int np_ = s_n;
  for (int k = 0; k < np_; ++k) {
    const HmPatch d = s_p[k];
    if (py < d.y0 || py >= d.y1 || px0 + VEC <= d.x0 || px0 >= d.x1) continue;
    const double dy = (double)(py - d.uly) - d.c0, dy2 = __dmul_rn(dy, dy);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int px = px0 + e;
      if (px >= d.x0 && px < d.x1) {
        const double dx = (double)(px - d.ulx) - d.c0;
        const double g = exp(-(__dadd_rn(__dmul_rn(dx, dx), dy2)) / d.den);
        v[e] = fmaxf(v[e], (float)g);                          // np.maximum into the float32 map
      }
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) v[e] = fminf(fmaxf(v[e], 0.0f), 1.0f);   // np.clip(target, 0, 1)
  float* dst = out + (size_t)plane * W * H + i;
  if (VEC == 4) *(float4*)dst = make_float4(v[0], v[1], v[2], v[3]);
  else
#pragma unroll
    for (int e = 0; e < VEC; ++e) dst[e] = v[e];
}

}  // namespace

size_t fvp_render_patch_bytes(int total_views, int max_people, int J) {
  return (size_t)total_views * max_people * J * sizeof(HmPatch);
}

void fvp_launch_render_heatmaps(const double* d_joints, const int* d_num, const unsigned char* d_vis, int total_views,
                                int max_people, int J, int W, int H, double stride_x, double stride_y, double sigma,
                                void* d_patches, float* d_out, cudaStream_t st) {
  const int n = total_views * max_people;
  k_hm_prepare<<<fvp_cdiv(n, 64), 64, 0, st>>>(d_joints, d_num, d_vis, total_views, max_people, J, W, H, stride_x, stride_y,
                                               sigma, (HmPatch*)d_patches);
  if (W % 4 == 0 && ((size_t)d_out & 15) == 0) {
    dim3 grid(total_views * J, fvp_cdiv(W * H, 1024));
    k_hm_render<4><<<grid, 256, 0, st>>>((const HmPatch*)d_patches, max_people, J, W, H, d_out);
  } else {
    dim3 grid(total_views * J, fvp_cdiv(W * H, 256));
    k_hm_render<1><<<grid, 256, 0, st>>>((const HmPatch*)d_patches, max_people, J, W, H, d_out);
  }
}
