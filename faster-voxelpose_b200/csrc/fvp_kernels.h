// fvp_kernels.h - host-side launchers of every kernel in libfvp_b200 (all enqueue on `st`).
#pragma once
#include "fvp_common.cuh"

// ---- geometry tables resident on the device ---------------------------------------------------
struct FvpGeom {
  FvpProj proj;
  int V, J, JG;              // views, joints, channel groups (JP/4)
  int X, Y, Z;               // coarse grid
  int fine[3];               // fine whole-space grid (project_individual.py:25)
  int P;                     // max people
  const float* coarse_axes;  // [X+Y+Z]
  const float* fine_axes;    // [fine0+fine1+fine2]
  const float* ind_axes;     // [64*3] individual-space axes incl. space centre (center_grid)
  const FvpSeq* seqs;        // [max_sequences]
  size_t view_stride4;       // float4 per view of the channel-last heat map (HP*WP*JG)
  // Sample-grid cache per calibration slot (the reference caches the same thing: project_whole.py:66-76,
  // project_individual.py sample_grid[seq]): unnormalised heat-map positions (ix, iy) of every grid voxel in every
  // view, computed ONCE per slot by fvp_project (k_build_sample_grid) and only looked up by K1 / K3.
  const float2* coarse_grid; // [max_sequences][X][Y][Z][V]
  const float2* fine_grid;   // [max_sequences][V][fine0][fine1][fine2]
};

// K0: [B][V][J][H][W] -> channel-last, zero-bordered [B][V][HP][WP][JP]
void fvp_launch_stage_heatmaps(const FvpGeom& g, const float* d_hm, float* d_hm_cl, int batch, cudaStream_t st);

// (re)build both sample grids of calibration slot `slot`
void fvp_launch_build_sample_grids(const FvpGeom& g, int slot, cudaStream_t st);

// K1: fused back-projection + view mean + clamp + z-max -> plane_cl [B][X][Y][JP]
void fvp_launch_hdn_project(const FvpGeom& g, const float* d_hm_cl, const int* d_frame_seq, float* d_plane_cl,
                            int batch, cudaStream_t st);

void fvp_launch_debug_project(const FvpGeom& g, int seq, const float* d_pts, int n, float* d_ix, float* d_iy,
                              cudaStream_t st);

// layout helpers (stage API / tests): NHWC(JP) <-> NCHW(J)
void fvp_launch_nhwc_to_nchw(const float* d_in, float* d_out, int n, int hw, int cp, int c, cudaStream_t st);
void fvp_launch_nchw_to_nhwc(const float* d_in, float* d_out, int n, int hw, int cp, int c, cudaStream_t st);

// K3: per-person back-projection + three-plane max -> planes_cl [3][B*P][64][64][JP] (zeroed, then RED.MAX-folded).
// A person is 64 (JG <= 4) or 128 column patches, times ncpart depth parts (1, 2, 4 or 8) - one CTA each.
static inline int fvp_k3_patches(int JG) { return JG <= 4 ? 64 : 128; }
void fvp_launch_jln_project(const FvpGeom& g, const float* d_hm_cl, const FvpPerson* d_people, float* d_planes_cl,
                            int batch, int ncpart, cudaStream_t st);

// N1 heat-map renderer (JointsDataset.generate_input_heatmap): joints [views][max_people][J][2] float64 in IMAGE_SIZE pixels,
// num [views], vis [views][max_people][J] or NULL -> out [views][J][H][W]; d_patches = fvp_render_patch_bytes(...) bytes
size_t fvp_render_patch_bytes(int total_views, int max_people, int J);
void fvp_launch_render_heatmaps(const double* d_joints, const int* d_num, const unsigned char* d_vis, int total_views,
                                int max_people, int J, int W, int H, double stride_x, double stride_y, double sigma,
                                void* d_patches, float* d_out, cudaStream_t st);

// ---- convolutions (fp32 CUDA-core implicit GEMM, NHWC) ----------------------------------------
struct FvpConvArgs {
  const float* in;    // [n][H][W][Cin]
  int H, W, Cin;
  const float* in2;   // optional second source for a fused 1x1 (skip_con), same H,W
  int Cin2;
  const float* w;     // packed [(taps*CinP + Cin2P)][CoutP], CinP = round_up(Cin,16)
  const float* bias;  // [CoutP]
  float* out;         // NHWC [n][Ho][Wo][CoutS]  or NCHW [n][CoutReal][Ho][Wo]
  int CoutP;          // GEMM N (multiple of 4); for upsample = 4*Co
  int CoutS;          // channel stride of the NHWC output
  int CoutReal;       // channels stored when nchw
  const float* res;   // residual, same layout as out (NHWC)
  int res_mode;       // 0 none, 1 add before ReLU, 2 add after ReLU
  int relu;
  int ksize;          // 1, 3, 7
  int upsample;       // ConvTranspose k2 s2 expressed as 1x1 conv to 4*Co + pixel shuffle
  int nchw;
  int n;
  const int* valid;   // optional per-image gate
  // Activation format (tensor-core engine 2 only; 0 = everything fp32 as described above).  A *split* tensor holds the
  // fp16 hi plane [n][H][W][C] followed by the fp16 scaled-lo plane (x = hi + lo * 2^-11, the operand form of the MMAs):
  // the producing epilogue splits once, consumers fetch halos with TMA tensor loads (no loader warps, no conversions).
  int fmt;            // FVP_FMT_* bits
};
enum { FVP_FMT_IN_SPLIT = 1, FVP_FMT_OUT_SPLIT = 2, FVP_FMT_RES_SPLIT = 4 };   // in / in2 share IN_SPLIT
void fvp_launch_conv(const FvpConvArgs& a, cudaStream_t st);
void fvp_launch_maxpool2(const float* in, float* out, int n, int H, int W, int C, const int* valid, cudaStream_t st);
// same on split tensors (hi/lo fp16 planes): the pair of the largest reconstructed value is copied, so the result is exact
void fvp_launch_maxpool2_split(const void* in, void* out, int n, int H, int W, int C, const int* valid, cudaStream_t st);
// element-wise fp32 <-> split conversion of `count` values (test hooks: a split tensor of `count` elements is 2*count halves)
void fvp_launch_split(const float* in, void* out, size_t count, cudaStream_t st);
void fvp_launch_unsplit(const void* in, float* out, size_t count, cudaStream_t st);

// ---- 2-D trunk program (CenterNet / P2PNet) ------------------------------------------------------
struct FvpConvW {         // one packed conv
  const float* w;         // [(taps*CinP + Cin2P)][CoutP] fp32 (CUDA-core kernel)
  const float* b;
  int cin, cin2, coutp, k;
  // weight images for the tcgen05 kernel, tiled per (K-block, tap, N-tile); index = N-tile cap: [0] <=128, [1] 32, [2] 64
  const float* wtc[3];      // tf32 hi/lo split (engine 1)
  const float* wtc16[3];    // fp16 hi / scaled-lo split (engine 2)
  const float* wtc16_c16;   // fp16 split with 16-channel K-blocks (layers with <= 16 input channels), NULL otherwise
};
// Everything a conv launch needs to know about its context.  Handed down from fvp_ctx on every call: the launchers keep NO
// process-global mutable state, so contexts on different devices / host threads do not interfere.
struct FvpLaunchEnv {
  int num_sms;                      // SM count of the context's device
  int conv_mode;                    // 0 CUDA cores, 1 tcgen05 3xTF32, 2 tcgen05 fp16 hi/lo split
  unsigned long long* tc_prof;      // debug: 12 role counters of k_conv_tc (fvp_debug_conv), NULL in production
  int* tc_plan;                     // host-only plan query (fvp_debug_conv_plan): decisions are recorded, nothing is launched
  int split_activations;            // engine 2: activations travel as split (fp16 hi / lo) tensors fetched by TMA (default on)
  int* error;                       // set to 1 when a launch could not be prepared (TMA descriptor encode failed); may be NULL
};
// tcgen05 / TMEM implicit-GEMM conv (fvp_conv_tc.cu); same arguments as fvp_launch_conv plus the tiled weights
void fvp_launch_conv_tc(const FvpConvArgs& a, const float* const wtc[3], int mode, const FvpLaunchEnv& env, cudaStream_t st);
void fvp_tc_geometry(int coutp, int narrow, int* n_tile, int* n_tiles);
// Per-DEVICE one-time setup (function attributes are per device: opt-in shared-memory sizes, carve-outs); fvp_create calls
// every one of these on the context's device, so a second context on another GPU of the same process is set up as well.
cudaError_t fvp_conv_tc_init_device();
cudaError_t fvp_conv_init_device();
cudaError_t fvp_proposal_init_device(int X, int Y);
// Range guard of the fp16 hi/lo engine: a conv output outside the fp16 range (the next layer's operand) sets *status |= 1
// through this per-device pointer (mapped pinned host memory, read by the host after a synchronise).
cudaError_t fvp_conv_tc_set_status_ptr(int* d_status);
cudaError_t fvp_conv_set_status_ptr(int* d_status);    // same word, CUDA-core kernel (its outputs may feed an fp16 layer)
struct FvpTrunkW {
  FvpConvW front, r1a, r1b, s1a, s1b, e1a, e1b, s2a, s2b, e2a, e2b, ma, mb, d2a, d2b, up2, d1a, d1b, up1;
  FvpConvW head_a, head_b;   // CenterNet: merged 3x3 (32->64) + block-diagonal 1x1 (64->3); P2PNet: head_b only
};
// buf[6]: scratch units of n*H*W*64 floats each
void fvp_run_trunk2d(const FvpTrunkW& t, const float* d_in, int cin, int n, int H, int W, float* const buf[6],
                     const int* valid, bool center_heads, float* d_out, int out_real, int* launches, cudaStream_t st,
                     const FvpLaunchEnv& env);

// ---- proposals -------------------------------------------------------------------------------
// nms2D + top-k: hm planar with image stride `img_stride` floats -> conf [B][P], flat [B][P]
void fvp_launch_nms_topk(const float* d_hm, size_t img_stride, int X, int Y, int P, int batch, float* d_conf,
                         int* d_flat, cudaStream_t st);

struct FvpC2CW {            // packed 1-D trunk, weights [tap][ci][co] per conv
  const float* w[24];
  const float* b[24];
  const float* w2[24];      // same, ci-major ([ci][tap][co], then the fused-skip rows) for the split-K kernel
  const void* plan;         // device copy of the weight-ring chunk table (fvp_c2c_build_plan)
  const float* w3;          // per-rank weight slices of the cluster kernel (fvp_c2c_pack_cluster); NULL = one CTA per column
  int max_clusters;         // clusters per launch of the cluster kernel (columns beyond that are looped over)
};
size_t fvp_c2c_cluster_floats(int J);
void fvp_c2c_pack_cluster(const float* const w2_host[20], const float* const b_host[20], int J, float* out);
int fvp_c2c_max_clusters(int J);       // co-resident 8-CTA clusters of the cluster kernel on the current device
size_t fvp_c2c_plan_bytes();
// chunk table of the network-wide weight ring of k_proposals from the DEVICE addresses of the ci-major weights; returns
// the number of chunks or < 0 (table too small / misaligned chunk)
int fvp_c2c_build_plan(const float* const w2[20], int J, void* h_plan);
struct FvpPropArgs {
  FvpGeom g;
  const float* hm_cl;
  const int* frame_seq;
  const float* conf2d;      // [B*P]
  const int* flat;          // [B*P]
  const float* size;        // planar: size_w at size[b*img_stride + (X*Y)*0 + flat], size_h at +X*Y
  size_t size_img_stride;
  const float* cols_in;     // optional [n][J][Z] columns (c2c stage API); else sampled from hm_cl
  float* cols_out;          // optional [n][J][Z]
  float* hm1d_out;          // optional [n][Z]
  float* centers;           // optional [B][P][7]
  FvpPerson* people;        // optional [B*P]
  int* img_valid;           // optional [3][n_slots]: gate of the three plane images of every slot
  int n_slots;
  float min_score;
  float hdn_scale[3], hdn_bias[3];                 // human_detection_net.py:22-23
  float jln_scale[3], jln_bias[3];                 // project_individual.py:27-29
  float whole[3], ind[3];
  int ind_vox[3];
  int mode;                 // 0 full (sample + c2c + proposal), 1 c2c only on cols_in
};
// prefer_latency: one 8-CTA cluster per column when the columns fit one wave of clusters (2.4x shorter at batch 1, ~2x the
// SM time); otherwise, and always for other Z than 20, one CTA per column
void fvp_launch_proposals(const FvpPropArgs& a, const FvpC2CW& w, int n, int prefer_latency, cudaStream_t st);
// proposal_centers [B][P][7] -> FvpPerson (stage API for K3)
void fvp_launch_people_from_centers(const FvpPropArgs& a, const float* d_centers, int n, cudaStream_t st);

// ---- pose head ----------------------------------------------------------------------------------
struct FvpPoseW {
  const float* conv_w;   // [32][9] BN-folded
  const float* conv_b;   // [32]
  const float* fc1_w;    // [hidden][feat]
  const float* fc1_b;
  const float* fc2_w;    // [hidden]
  const float* fc2_b;    // [1]
  int feat, hidden;
};
// feat planar [3][n][J][64][64]; people may be NULL (then offsets from d_offset [n][3], all valid)
void fvp_launch_pose_head(const FvpGeom& g, const FvpPoseW& w, const float* d_feat, const FvpPerson* d_people,
                          const float* d_offset, int n, float beta, float* d_pose, float* d_maxw, float* d_weights,
                          float* d_fused, cudaStream_t st);
// conf[n] = mean over (plane, joint) of max softmax; writes fused_poses/plane_poses/centers col 4 (final pack)
void fvp_launch_finalize(const FvpGeom& g, const FvpPerson* d_people, const float* d_maxw, const float* d_pose,
                         const float* d_fused, float* d_centers, int batch, float* d_conf, float* d_fused_poses,
                         float* d_plane_poses, float* d_centers_out, cudaStream_t st);
