// fvp_backproject.cu - the fused back-projection kernels.
//
//   K0  stage heat maps : [B][V][J][H][W] -> channel-last [B][V][HP][WP][JP] with a zero border wide
//                         enough for every tap of a clamped (|g|<=1.1) sample -> no bounds tests later.
//   K1  HDN            : ProjectLayer(whole).forward (project_whole.py:62-88) + CenterNet's z-max
//                         (cnns_2d.py:174): plane[b,j,x,y] = max_z clamp(mean_v bilinear(hm[b,v,j], proj_v(x,y,z)))
//                         The [B,J,X,Y,Z] volume never exists.
//   K3  JLN            : ProjectLayer(individual).forward (project_individual.py:96-136) + the three
//                         orthographic max planes (joint_localization_net.py:80-81).  The [N,J,64^3] cubes
//                         never exist.
//   Sample positions: K1 / K3 look up per-calibration sample-grid caches (k_build_sample_grid, the same bit-exact
//   fvp_project chain evaluated once per calibration slot, like the reference's own per-sequence grids).  Measured on
//   B200 in round 2 (profiles/r02_k3_experiments.txt): recomputing the projection inside K3 cuts its DRAM traffic to
//   the algorithmic bytes but makes it 20 % slower (0.135 -> 0.162 ms at batch 1, 3.50 -> 4.22 ms at batch 32) - the
//   kernel is bound by the L1 gather and its issue slots, not by HBM - so the cache stays and that variant was removed.
//
// Lane layout of K1/K3: CG consecutive lanes own one voxel column, lane s of the group holds channel
// group s (4 joints, one float4).  One tap of one voxel is therefore a single contiguous 16*JG byte
// record and a warp-wide LDG.128 touches 32/CG records -> fully used sectors.  The CG lanes of a column
// split the (z, view) projection work between them and exchange tap descriptors by shuffle.
#include "fvp_kernels.h"
#include <cstdlib>
#include "fvp_project.cuh"

// ------------------------------------------------------------------------------------------------
// K0
// ------------------------------------------------------------------------------------------------
template <int JG>
__global__ void __launch_bounds__(256) k0_stage_heatmaps(const float* __restrict__ hm, float4* __restrict__ out,
                                                          int J, int H, int W, int WP, int PADX, int PADY,
                                                          size_t view_stride4) {
  const int bv = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= H * W) return;
  const float* src = hm + (size_t)bv * J * H * W + pix;
  float v[JG * 4];
#pragma unroll
  for (int j = 0; j < JG * 4; ++j) v[j] = (j < J) ? __ldg(src + (size_t)j * H * W) : 0.0f;
  const int y = pix / W, x = pix - y * W;
  float4* dst = out + (size_t)bv * view_stride4 + ((size_t)(y + PADY) * WP + (x + PADX)) * JG;
#pragma unroll
  for (int g = 0; g < JG; ++g) dst[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}

void fvp_launch_stage_heatmaps(const FvpGeom& g, const float* d_hm, float* d_hm_cl, int batch, cudaStream_t st) {
  const FvpProj& P = g.proj;
  dim3 grid(fvp_cdiv(P.H * P.W, 256), batch * g.V);
  if (g.JG == 4)
    k0_stage_heatmaps<4><<<grid, 256, 0, st>>>(d_hm, (float4*)d_hm_cl, g.J, P.H, P.W, P.WP, P.PADX, P.PADY, g.view_stride4);
  else
    k0_stage_heatmaps<5><<<grid, 256, 0, st>>>(d_hm, (float4*)d_hm_cl, g.J, P.H, P.W, P.WP, P.PADX, P.PADY, g.view_stride4);
}

// ------------------------------------------------------------------------------------------------
// sample-grid cache: (ix, iy) of every coarse / fine grid voxel in every view of one calibration slot
// ------------------------------------------------------------------------------------------------
// fine != 0: out[v][x][y][z] over the fine grid; fine == 0: out[x][y][z][v] over the coarse grid (K1 walks (z, view)
// pairs of a column, which are contiguous in that order).  Same fvp_project(), same axis values as the kernels used
// to evaluate in place, so the cached positions are bit-identical to the on-the-fly ones.
__global__ void __launch_bounds__(256) k_build_sample_grid(FvpGeom g, int slot, int fine, float2* __restrict__ out) {
  const int n0 = fine ? g.fine[0] : g.X, n1 = fine ? g.fine[1] : g.Y, n2 = fine ? g.fine[2] : g.Z;
  const float* ax = fine ? g.fine_axes : g.coarse_axes;
  const int nvox = n0 * n1 * n2;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nvox * g.V) return;
  int v, vox;
  if (fine) { v = i / nvox; vox = i - v * nvox; } else { vox = i / g.V; v = i - vox * g.V; }
  const int z = vox % n2, xy = vox / n2, y = xy % n1, x = xy / n1;
  const FvpSeq& sq = g.seqs[slot];
  float ix, iy;
  fvp_project(sq.cam[v], sq.A, g.proj, ax[x], ax[n0 + y], ax[n0 + n1 + z], ix, iy);
  out[i] = make_float2(ix, iy);
}
void fvp_launch_build_sample_grids(const FvpGeom& g, int slot, cudaStream_t st) {
  const size_t nc = (size_t)g.X * g.Y * g.Z * g.V, nf = (size_t)g.fine[0] * g.fine[1] * g.fine[2] * g.V;
  k_build_sample_grid<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(g, slot, 0, const_cast<float2*>(g.coarse_grid) + slot * nc);
  k_build_sample_grid<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(g, slot, 1, const_cast<float2*>(g.fine_grid) + slot * nf);
}

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
// CTA = 4 warps that all own the same 32/CG voxel columns; warp w takes the z range [w*Z/4, (w+1)*Z/4) so a
// batch-1 launch still puts ~5 CTAs on every SM; the four partial z-maxima meet in shared memory.
template <int CG>
__global__ void __launch_bounds__(128) k1_hdn_project_zmax(FvpGeom g, const float4* __restrict__ hm_cl,
                                                            const int* __restrict__ frame_seq,
                                                            float4* __restrict__ plane_cl) {
  __shared__ float4 s_part[4][32];
  const int b = blockIdx.y;
  const FvpProj& P = g.proj;

  constexpr int COLS_PER_WARP = 32 / CG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % CG;                       // channel group / projection sub-lane
  const int group_base = lane - s;               // first lane of my column
  const int ncols = g.X * g.Y;
  int col = blockIdx.x * COLS_PER_WARP + lane / CG;
  const bool col_ok = col < ncols;
  if (!col_ok) col = ncols - 1;                  // keep the warp convergent for the shuffles
  // cached sample positions of this column: [z][view] pairs, contiguous
  const float2* grid_col = g.coarse_grid + ((size_t)frame_seq[b] * g.X * g.Y + col) * g.Z * g.V;

  const int V = g.V, ZW = g.Z >> 2, z0 = warp * ZW, npairs = ZW * V;
  const float fV = (float)V, rV = 1.0f / fV;
  const int row4 = P.WP * g.JG, px4 = g.JG;
  const bool ch_ok = s < g.JG;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), zmax = acc;
  for (int base = 0; base < npairs; base += CG) {
    const int idx = base + s;                    // my (z, view) pair of this round
    FvpTaps t;
    t.off = 0; t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
    if (idx < npairs) {
      const int z = idx / V, v = idx - z * V;
      const float2 q = __ldg(grid_col + z0 * V + idx);
      t = fvp_taps(P, q.x, q.y);
      t.off += (b * V + v) * (int)g.view_stride4;
    }
#pragma unroll
    for (int k = 0; k < CG; ++k) {
      const int ik = base + k;
      if (ik >= npairs) break;                   // warp-uniform
      const int off = __shfl_sync(0xffffffffu, t.off, group_base + k);
      const float w00 = __shfl_sync(0xffffffffu, t.w00, group_base + k);
      const float w01 = __shfl_sync(0xffffffffu, t.w01, group_base + k);
      const float w10 = __shfl_sync(0xffffffffu, t.w10, group_base + k);
      const float w11 = __shfl_sync(0xffffffffu, t.w11, group_base + k);
      const int zk = ik / V, vk = ik - zk * V;
      if (ch_ok) fvp_tap_accumulate(acc, hm_cl, off + s, row4, px4, w00, w01, w10, w11);
      if (vk == V - 1) {                         // last view of this z: mean, clamp, running z-max
        zmax = fvp_max4(zmax, fvp_mean_clamp4(acc, fV, rV));
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  s_part[warp][lane] = zmax;
  __syncthreads();
  if (warp == 0 && col_ok && ch_ok) {
    const float4 m = fvp_max4(fvp_max4(s_part[0][lane], s_part[1][lane]), fvp_max4(s_part[2][lane], s_part[3][lane]));
    plane_cl[((size_t)b * ncols + col) * g.JG + s] = m;
  }
}

void fvp_launch_hdn_project(const FvpGeom& g, const float* d_hm_cl, const int* d_frame_seq, float* d_plane_cl,
                            int batch, cudaStream_t st) {
  const int ncols = g.X * g.Y;
  const size_t smem = 0;
  if (g.JG <= 4) {
    dim3 grid(fvp_cdiv(ncols, 8), batch);
    k1_hdn_project_zmax<4><<<grid, 128, smem, st>>>(g, (const float4*)d_hm_cl, d_frame_seq, (float4*)d_plane_cl);
  } else {
    dim3 grid(fvp_cdiv(ncols, 4), batch);
    k1_hdn_project_zmax<8><<<grid, 128, smem, st>>>(g, (const float4*)d_hm_cl, d_frame_seq, (float4*)d_plane_cl);
  }
}

// ------------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_nhwc_to_nchw(const float* __restrict__ in, float* __restrict__ out, int hw, int cp, int c) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw * c) return;
  const int ch = i / hw, p = i - ch * hw;
  out[(size_t)n * c * hw + i] = in[((size_t)n * hw + p) * cp + ch];
}
__global__ void k_nchw_to_nhwc(const float* __restrict__ in, float* __restrict__ out, int hw, int cp, int c) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw * cp) return;
  const int p = i / cp, ch = i - p * cp;
  out[(size_t)n * hw * cp + i] = ch < c ? in[((size_t)n * c + ch) * hw + p] : 0.0f;
}
void fvp_launch_nhwc_to_nchw(const float* d_in, float* d_out, int n, int hw, int cp, int c, cudaStream_t st) {
  dim3 grid(fvp_cdiv(hw * c, 256), n);
  k_nhwc_to_nchw<<<grid, 256, 0, st>>>(d_in, d_out, hw, cp, c);
}
void fvp_launch_nchw_to_nhwc(const float* d_in, float* d_out, int n, int hw, int cp, int c, cudaStream_t st) {
  dim3 grid(fvp_cdiv(hw * cp, 256), n);
  k_nchw_to_nhwc<<<grid, 256, 0, st>>>(d_in, d_out, hw, cp, c);
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------
// max-fold a float4 of values >= +0 into a zero-initialised plane: one RED.MAX.U32 per non-zero component
// (x > 0 is false for +0, -0 and NaN, none of which may enter an unsigned comparison)
__device__ __forceinline__ void fvp_red_max4(float4* dst, float4 m) {
  unsigned* p = reinterpret_cast<unsigned*>(dst);
  if (m.x > 0.0f) atomicMax(p + 0, __float_as_uint(m.x));
  if (m.y > 0.0f) atomicMax(p + 1, __float_as_uint(m.y));
  if (m.z > 0.0f) atomicMax(p + 2, __float_as_uint(m.z));
  if (m.w > 0.0f) atomicMax(p + 3, __float_as_uint(m.w));
}


// ------------------------------------------------------------------------------------------------
// One CTA = one person x one compact patch of 8 cube rows a (one per warp) x BPW = 32/CG cube columns b x one depth
// part.  A pass of the patch over one view and one chunk of CG depths touches a compact ~(8*1.5)^2-pixel window of the
// heat map (a 64-column sheet of one row, the first design, touched ~4x more pixels per pass), so the 2x2 footprints of neighbouring voxels share L1 lines whatever
// the direction the camera looks along.  No shared state survives a chunk:
//   xy[a][b] = max_c : the thread owns (a, b) - a register, stored once (complete when there is one depth part)
//   xz[a][c] = max_b : the BPW columns of a row meet in shared memory; partial over the 64/BPW b-blocks
//   yz[b][c] = max_a : the 8 warps (8 rows) meet in the same shared-memory image; partial over the 8 a-blocks
// Partial maxima are folded straight into the zero-initialised output planes with RED.MAX on the float bit patterns
// (values are >= +0, so they order like unsigned ints; max is exact and order-independent, the result is deterministic).
// Only non-zero values are sent - heat maps are sparse, >90 % of the partial maxima are 0 - so there are no scratch
// images and no second reduce kernel: DRAM traffic stays near the algorithmic bytes.
// XCH: how the lane that looked up a depth hands its taps to the other lanes of its column group.  Shuffles ride the LSU
// pipe that bounds this kernel, so fewer of them is time (measured on B200, profiles/r02_k3_xch.txt):
//   0 = five shuffles per depth and view (offset + four weights);
//   1 = three shuffles (offset + the two fractions), the weights rebuilt per lane with the very operations of fvp_taps
//       (bit-identical): K3 3.46 -> 3.35 ms at batch 32, 0.125 -> 0.121 ms at batch 1.  Used for 64-byte records
//       (JG == 4), the instantiation it was measured and parity-tested on.
// (A third form - exchange through the idle shared-memory chunk image, one LDS.128 per depth - measured 3.58 ms: removed.)
template <int CG, int MINB, int PX16, int XCH = 0>
__global__ void __launch_bounds__(256, MINB)
k3_jln_patch(FvpGeom g, const float4* __restrict__ hm_cl, const FvpPerson* __restrict__ people,
             float4* __restrict__ planes_cl, int n_people, int ncpart) {
  constexpr int BPW = 32 / CG;                   // cube columns b per warp = patch extent in b
  constexpr int NBB = 64 / BPW;                  // b-blocks per cube
  constexpr int NAB = 8;                         // a-blocks per cube (8 rows each, one row per warp)
  constexpr int CCH = CG;                        // depths per chunk: one projection round of the lane group
  constexpr int NBUF = (CG == 4) ? 2 : 1;        // double buffer when it fits the 48 KB static limit
  __shared__ float4 s_yz[NBUF][CCH][8][32];      // [depth][warp = row][lane = (b, channel group)]

  const int person = blockIdx.y;
  const int patch = blockIdx.x % (NAB * NBB), cpart = blockIdx.x / (NAB * NBB);
  const int ablk = patch / NBB, bblk = patch - ablk * NBB;
  const int c_begin = cpart * (64 / ncpart), c_end = c_begin + 64 / ncpart;
  const FvpPerson pd = people[person];
  const FvpProj& P = g.proj;
  const int JG = g.JG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = lane % CG, group_base = lane - s;
  const int a = ablk * 8 + warp;                 // cube row    (world x index within the cube), warp-uniform
  const int b = bblk * BPW + lane / CG;          // cube column (world y index within the cube)
  const bool ch_ok = s < JG;
  const size_t img4 = (size_t)64 * 64 * JG;      // float4 per plane image
  float4* xy_img = planes_cl + ((size_t)0 * n_people + person) * img4;
  float4* xz_img = planes_cl + ((size_t)1 * n_people + person) * img4;
  float4* yz_img = planes_cl + ((size_t)2 * n_people + person) * img4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  const bool live = pd.valid && !pd.empty;
  const bool any = live && max(ablk * 8, pd.lo[0]) < min(ablk * 8 + 8, pd.hi[0]) &&
                   max(bblk * BPW, pd.lo[1]) < min(bblk * BPW + BPW, pd.hi[1]) &&
                   max(c_begin, pd.lo[2]) < min(c_end, pd.hi[2]);
  if (!any) return;                              // nothing to sample in this patch: the planes are already zero

  const bool a_ok = a >= pd.lo[0] && a < pd.hi[0];           // warp-uniform
  const bool b_ok = b >= pd.lo[1] && b < pd.hi[1];
  const int F1 = g.fine[1], F2 = g.fine[2], nfine = g.fine[0] * F1 * F2;
  const float2* grid_s = g.fine_grid + (size_t)pd.seq * nfine * g.V;
  const int col = ((pd.tl[0] + a) * F1 + pd.tl[1] + b) * F2 + pd.tl[2];     // + view * nfine + c
  const int V = g.V;
  const float fV = (float)V, rV = 1.0f / fV;
  const int row4 = P.WP * JG, px4 = JG;
  const int vs4 = (int)g.view_stride4, frame_off = (person / g.P) * V * vs4;
  const bool sample_ok = ch_ok && b_ok && a_ok;
  const int s_me = ch_ok ? s : JG - 1;           // idle channel-group lanes re-read the last group (same sectors)

  float4 xy_m = zero4;
  int it = 0;
  for (int cc = c_begin; cc < c_end; cc += CCH, ++it) {
    float4 acc[CCH];
#pragma unroll
    for (int c = 0; c < CCH; ++c) acc[c] = zero4;
    // Every branch of the sampling loop is a vote, i.e. provably warp-uniform: lanes outside the crop (b, channel group)
    // sample pixel (0,0) of the view like everybody else and are zeroed afterwards - no divergent regions.
    const bool row_live = __any_sync(0xffffffffu, a_ok && max(cc, pd.lo[2]) < min(cc + CCH, pd.hi[2]));
    if (row_live) {
      const int cz = cc + s;                     // the depth this lane looks up for its column
      const bool q_ok = b_ok && cz >= pd.lo[2] && cz < pd.hi[2];
      // The next view's cached position is fetched one view early (ncu attributed 9 % of K3's stall samples to the first
      // use of this load; measured +1 % at batch 32, profiles/r02_k3_experiments.txt).
      float2 q_next = make_float2(0.f, 0.f);
      if (q_ok) q_next = __ldg(grid_s + col + cz);
      for (int v = 0; v < V; ++v) {
        FvpTapRegs tr;
        tr.off = -1;
        tr.a = tr.b = tr.c = tr.d = zero4;
        const float2 q = q_next;
        if (q_ok && v + 1 < V) q_next = __ldg(grid_s + (size_t)(v + 1) * nfine + col + cz);
        if (XCH == 0) {
          const FvpTaps t = fvp_taps(P, q.x, q.y);
          const int my_off = t.off + frame_off + v * vs4;
#pragma unroll
          for (int k = 0; k < CG; ++k) {         // compile-time depth index within the chunk
            const int off = __shfl_sync(0xffffffffu, my_off, group_base + k);
            const float w00 = __shfl_sync(0xffffffffu, t.w00, group_base + k);
            const float w01 = __shfl_sync(0xffffffffu, t.w01, group_base + k);
            const float w10 = __shfl_sync(0xffffffffu, t.w10, group_base + k);
            const float w11 = __shfl_sync(0xffffffffu, t.w11, group_base + k);
            if (__any_sync(0xffffffffu, (cc + k) >= pd.lo[2] && (cc + k) < pd.hi[2]))
              fvp_tap_accumulate_vote<PX16>(acc[k], tr, hm_cl, off + s_me, row4, px4, w00, w01, w10, w11);
          }
        } else {
          int my_off;
          float my_fx, my_fy;
          fvp_taps_frac(P, q.x, q.y, my_off, my_fx, my_fy);
          my_off += frame_off + v * vs4;
#pragma unroll
          for (int k = 0; k < CG; ++k) {
            const int off = __shfl_sync(0xffffffffu, my_off, group_base + k);
            const float fx = __shfl_sync(0xffffffffu, my_fx, group_base + k);
            const float fy = __shfl_sync(0xffffffffu, my_fy, group_base + k);
            float w00, w01, w10, w11;
            fvp_tap_weights(fx, fy, w00, w01, w10, w11);
            if (__any_sync(0xffffffffu, (cc + k) >= pd.lo[2] && (cc + k) < pd.hi[2]))
              fvp_tap_accumulate_vote<PX16>(acc[k], tr, hm_cl, off + s_me, row4, px4, w00, w01, w10, w11);
          }
        }
      }
    }
    // mean + clamp + the three maxima of this chunk.  Both cross-thread maxima go through ONE shared-memory image of
    // the chunk ([depth][row][lane]): a partial-mask REDUX per channel group costs a serialised collective per mask
    // (CREDUX + ENDCOLLECTIVE + BSSY/BSYNC were ~20 % of the instructions and ~35 % of the stall samples in ncu).
    float4(*yzb)[8][32] = s_yz[NBUF == 2 ? (it & 1) : 0];
#pragma unroll
    for (int c = 0; c < CCH; ++c) {
      const bool c_in = (cc + c) >= pd.lo[2] && (cc + c) < pd.hi[2];
      const float4 val = (sample_ok && c_in) ? fvp_mean_clamp4(acc[c], fV, rV) : zero4;   // outside the crop: 0
      xy_m = fvp_max4(xy_m, val);
      yzb[c][warp][lane] = val;
    }
    __syncthreads();
    constexpr int N_YZ = CCH * 32;               // yz outputs of the chunk: (depth, b, channel group)
    constexpr int N_XZ = 8 * CCH * CG;           // xz outputs of the chunk: (row, depth, channel group)
    for (int o = tid; o < N_YZ + N_XZ; o += 256) {
      if (o < N_YZ) {                            // yz[b][cc + c] = max over the 8 rows of the patch
        const int c = o >> 5, l = o & 31, ss = l % CG;
        if (ss < JG) {
          float4 m = yzb[c][0][l];
#pragma unroll
          for (int w = 1; w < 8; ++w) m = fvp_max4(m, yzb[c][w][l]);
          fvp_red_max4(yz_img + ((size_t)(bblk * BPW + l / CG) * 64 + cc + c) * JG + ss, m);
        }
      } else {                                   // xz[a][cc + c] = max over the BPW columns of the patch
        const int x = o - N_YZ, ss = x % CG, c = (x / CG) % CCH, w = x / (CG * CCH);
        if (ss < JG) {
          float4 m = yzb[c][w][ss];
#pragma unroll
          for (int i = 1; i < BPW; ++i) m = fvp_max4(m, yzb[c][w][i * CG + ss]);
          fvp_red_max4(xz_img + ((size_t)(ablk * 8 + w) * 64 + cc + c) * JG + ss, m);
        }
      }
    }
    if (NBUF == 1) __syncthreads();
  }
  if (ch_ok) {
    if (ncpart == 1) xy_img[((size_t)a * 64 + b) * JG + s] = xy_m;           // complete: plain store
    else fvp_red_max4(xy_img + ((size_t)a * 64 + b) * JG + s, xy_m);          // partial over the depth parts
  }
}

void fvp_launch_jln_project(const FvpGeom& g, const float* d_hm_cl, const FvpPerson* d_people, float* d_planes_cl,
                            int batch, int ncpart, cudaStream_t st) {
  const int n_people = batch * g.P;
  const int img4 = 64 * 64 * g.JG;
  // partial maxima are RED-folded into the planes, which therefore start from zero
  cudaMemsetAsync(d_planes_cl, 0, (size_t)3 * n_people * img4 * sizeof(float4), st);
  dim3 grid(fvp_k3_patches(g.JG) * ncpart, n_people);
  if (g.JG == 4)
    k3_jln_patch<4, 4, 64, 1><<<grid, 256, 0, st>>>(g, (const float4*)d_hm_cl, d_people, (float4*)d_planes_cl, n_people, ncpart);
  else if (g.JG < 4)
    k3_jln_patch<4, 4, 0><<<grid, 256, 0, st>>>(g, (const float4*)d_hm_cl, d_people, (float4*)d_planes_cl, n_people, ncpart);
  else
    k3_jln_patch<8, 2, 0><<<grid, 256, 0, st>>>(g, (const float4*)d_hm_cl, d_people, (float4*)d_planes_cl, n_people, ncpart);
}

// ------------------------------------------------------------------------------------------------
// test hook: the in-kernel projection chain on arbitrary world points (bit-parity test vs project_chain_np)
// ------------------------------------------------------------------------------------------------
__global__ void k_debug_project(FvpGeom g, int seq, const float* __restrict__ pts, int n, float* __restrict__ ix,
                                float* __restrict__ iy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (i >= n) return;
  const FvpSeq& sq = g.seqs[seq];
  float x, y;
  fvp_project(sq.cam[v], sq.A, g.proj, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y);
  ix[(size_t)v * n + i] = x;
  iy[(size_t)v * n + i] = y;
}
void fvp_launch_debug_project(const FvpGeom& g, int seq, const float* d_pts, int n, float* d_ix, float* d_iy,
                              cudaStream_t st) {
  dim3 grid(fvp_cdiv(n, 128), g.V);
  k_debug_project<<<grid, 128, 0, st>>>(g, seq, d_pts, n, d_ix, d_iy);
}
