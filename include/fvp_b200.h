/*
 * fvp_b200.h - C ABI of libfvp_b200.so: the B200-native (sm_100a) inference hot path of
 * Faster-VoxelPose (AlvinYH/Faster-VoxelPose @ 733aa98).
 *
 * The reference is pure Python/PyTorch and has no FFI; the native boundary it *would* bind is the
 * set of nn.Module.forward() calls on the hot path.  Each entry point below
 * names the reference interface it replaces.  All pointers are plain host or device pointers, all
 * sizes are explicit, no torch types appear.  Every function returns 0 on success or a negative
 * FVP_E_* code; fvp_last_error() gives the message (the Python host layer turns it into the
 * exception the reference would have raised: AssertionError for calibration mismatches,
 * RuntimeError otherwise).
 *
 * Tensor layouts are the reference's (row-major, fp32) unless a name ends in _cl (channel-last,
 * channels padded to a multiple of 4).  "d_" = device pointer, "h_" = host pointer.
 * All device work is enqueued on the cudaStream_t passed as `stream` (a CUstream handle as
 * uintptr; 0 = legacy default stream) and is CUDA-graph capturable unless stated otherwise.
 */
#ifndef FVP_B200_H_
#define FVP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVP_ABI_VERSION 1

enum {
  FVP_OK = 0,
  FVP_E_INVALID = -1,   /* bad argument / shape                                   */
  FVP_E_CUDA = -2,      /* CUDA runtime error (message has cudaGetErrorString)    */
  FVP_E_STATE = -3,     /* call order (e.g. forward before finalize_params)       */
  FVP_E_NOTFOUND = -4,  /* unknown parameter / sequence slot                      */
  FVP_E_CALIB = -5,     /* calibration missing / wrong camera count (reference: AssertionError,
                           lib/models/project_whole.py:73-74)                     */
  FVP_E_RANGE = -6      /* an activation left the fp16 range of the hi/lo tensor-core engine (results are invalid;
                           select conv mode 1 or 0).  Reported by the host entry points and fvp_check_range.   */
};

/* Geometry + network constants: the keys FasterVoxelPoseNet reads from cfg (SURVEY.md section 5,
 * lib/models/project_whole.py:16-23, project_individual.py:17-33, human_detection_net.py:19-23,
 * joint_localization_net.py:18, weight_net.py:51-54). */
typedef struct fvp_config {
  int32_t num_views;          /* DATASET.CAMERA_NUM                         */
  int32_t num_joints;         /* DATASET.NUM_JOINTS                         */
  int32_t hm_w, hm_h;         /* DATASET.HEATMAP_SIZE  (w, h)               */
  float image_w, image_h;     /* DATASET.IMAGE_SIZE                         */
  float ori_w, ori_h;         /* DATASET.ORI_IMAGE_SIZE                     */
  float space_size[3];        /* CAPTURE_SPEC.SPACE_SIZE                    */
  float space_center[3];      /* CAPTURE_SPEC.SPACE_CENTER                  */
  int32_t voxels[3];          /* CAPTURE_SPEC.VOXELS_PER_AXIS               */
  float ind_space_size[3];    /* INDIVIDUAL_SPEC.SPACE_SIZE                 */
  int32_t ind_voxels[3];      /* INDIVIDUAL_SPEC.VOXELS_PER_AXIS (64,64,64) */
  int32_t max_people;         /* CAPTURE_SPEC.MAX_PEOPLE                    */
  float min_score;            /* CAPTURE_SPEC.MIN_SCORE                     */
  float beta;                 /* NETWORK.BETA                               */
  int32_t feat_channels;      /* NETWORK.NUM_CHANNEL_JOINT_FEAT   (32)      */
  int32_t hidden_channels;    /* NETWORK.NUM_CHANNEL_JOINT_HIDDEN (64)      */
  int32_t max_batch;          /* frames per forward call the workspaces are sized for */
  int32_t max_sequences;      /* distinct calibrations kept resident        */
} fvp_config;

typedef struct fvp_ctx fvp_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* replaces models.faster_voxelpose.get(cfg) (lib/models/faster_voxelpose.py:108-110) */
int fvp_create(const fvp_config* cfg, int device, fvp_ctx** out);
/* A lane of `root`: a context for one more frame in flight on the same GPU.  It owns workspaces, streams and a CUDA graph
 * (sized for max_batch frames) and shares root's weights, axis tables, calibrations and sample-grid caches, so L frames in
 * flight cost L workspaces but ONE weight set and ONE grid cache.  Parameters, axes and calibrations are set on the root
 * only (FVP_E_STATE on a lane) and are picked up by the lanes at their next forward; the caller must not change them
 * while forwards of any lane are in flight.  Destroy lanes before their root (fvp_destroy(root) refuses otherwise -
 * it then returns without freeing and fvp_last_error(root) explains). */
int fvp_create_lane(fvp_ctx* root, int max_batch, fvp_ctx** out);
/* end of life of the model object (the reference relies on Python garbage collection); frees device memory */
void fvp_destroy(fvp_ctx* ctx);
const char* fvp_last_error(const fvp_ctx* ctx);   /* ctx may be NULL: error of the last failed fvp_create */
int fvp_abi_version(void);

/* ---- parameters: replaces nn.Module.load_state_dict (run/validate.py:78-81) ------------------ */
/* The table lists the reference's 485 state_dict keys in order (SURVEY.md section 5 "checkpoint"). */
int fvp_param_count(const fvp_ctx* ctx);
const char* fvp_param_name(const fvp_ctx* ctx, int index);
int64_t fvp_param_numel(const fvp_ctx* ctx, int index);
/* replaces model.load_state_dict(torch.load(model_file)) (run/validate.py:78-81), one tensor at a time:
 * copy one fp32 state_dict tensor (host memory, contiguous, reference shape); the int64
 * num_batches_tracked entries are accepted and ignored (numel 1, pass NULL or anything). */
int fvp_set_param(fvp_ctx* ctx, const char* name, const float* h_data, int64_t numel);
/* fold eval-mode BatchNorm (eps 1e-5) into the convolutions, repack for the kernels, upload.
 * Must be called after all parameters were set and again after any later fvp_set_param. */
int fvp_finalize_params(fvp_ctx* ctx);

/* ---- geometry tables (optional) ---------------------------------------------------------------
 * The reference builds voxel coordinates with torch.linspace (project_whole.py:34-40), whose fp32
 * values are not exactly start+i*step.  By default the library uses the symmetric scalar formula of
 * ATen's CUDA linspace; a host that wants bit-parity with a particular CPU run passes the exact
 * tables: coarse axes (X+Y+Z floats), fine axes (fineX+fineY+fineZ floats, project_individual.py:41)
 * and individual axes (64+64+64 floats: the center_grid of project_individual.py:37-40). */
int fvp_set_axes(fvp_ctx* ctx, const float* h_coarse, const float* h_fine, const float* h_individual);
int fvp_fine_voxels(const fvp_ctx* ctx, int32_t out[3]);

/* ---- calibration: replaces the per-sequence sample-grid caches (project_whole.py:75-80,
 *      project_individual.py:104-106) by a 21-float camera block per view ---------------------- */
/* h_cameras: [num_views][21] fp32 = R(9 row-major) T(3) fx fy cx cy k(3) p(2)
 * h_resize : [6] fp32 resize_transform (2x3 row-major)                       */
int fvp_set_sequence(fvp_ctx* ctx, int slot, const float* h_cameras, int num_views, const float* h_resize);

/* ---- whole forward: replaces FasterVoxelPoseNet.forward, eval branch
 *      (lib/models/faster_voxelpose.py:34-48,99-105) ------------------------------------------- */
/* d_heatmaps        [batch][V][J][H][W] fp32 (input_heatmaps)
 * h_seq_slots       [batch] calibration slot of every frame (meta['seq'] -> slot)
 * d_fused_poses     [batch][P][J][5]   (x,y,z mm, flag, conf)
 * d_plane_poses     [3][batch][P][J][2]
 * d_proposal_centers[batch][P][7]
 * Any output pointer may be NULL. */
int fvp_forward(fvp_ctx* ctx, const float* d_heatmaps, int batch, const int32_t* h_seq_slots,
                float* d_fused_poses, float* d_plane_poses, float* d_proposal_centers, uintptr_t stream);

/* One iteration of the reference's validation loop (run/validate.py:95-105: input_heatmaps.to(DEVICE), model(...), and
 * the later read-back of the poses) as one call with HOST buffers (pinned or pageable): H2D of the heatmaps, forward, D2H of the
 * outputs, all on `stream`, followed by a stream synchronise.  This is the end-to-end entry the
 * benchmark's `e2e` figure times. */
int fvp_forward_host(fvp_ctx* ctx, const float* h_heatmaps, int batch, const int32_t* h_seq_slots,
                     float* h_fused_poses, float* h_plane_poses, float* h_proposal_centers, uintptr_t stream);

/* Pipelined form of fvp_forward_host for a stream of frames (the reference's validate loop, lib/core/function.py:
 * one model(...) call per data-loader batch): fvp_submit_host enqueues H2D -> forward -> D2H on the context's own
 * streams and returns a ticket at once; the H2D of the next step overlaps the kernels of the current one (two device
 * input buffers).  Results are in the host buffers once fvp_wait(ticket) returns.  At most two tickets may be
 * outstanding (FVP_E_STATE otherwise); tickets complete in order; host buffers of an outstanding ticket must not be
 * reused.  Pinned host memory is required for real overlap, pageable memory still works. */
int fvp_submit_host(fvp_ctx* ctx, const float* h_heatmaps, int batch, const int32_t* h_seq_slots,
                    float* h_fused_poses, float* h_plane_poses, float* h_proposal_centers, long long* ticket);
int fvp_wait(fvp_ctx* ctx, long long ticket);

/* Capture the forward for (batch, seq slots) into a CUDA graph bound to fixed internal I/O buffers
 * and replay it on later fvp_forward* calls with the same signature (1 = on, 0 = off). */
int fvp_use_cuda_graph(fvp_ctx* ctx, int enable);

/* ---- stage entry points (parity tests + profiling; same kernels fvp_forward launches) -------- */
/* test hook: the in-kernel voxel -> heat-map-pixel chain (project_whole.py:49-60 + grid_sample un-normalise) on
 * arbitrary world points d_points [n][3] -> d_ix, d_iy [V][n]; bit-identical to oracle.project_chain_np */
int fvp_debug_project(fvp_ctx* ctx, int slot, const float* d_points, int n, float* d_ix, float* d_iy, uintptr_t stream);
/* test hook: one standalone NHWC convolution (d_in [n][H][W][cin], cin % 4 == 0; h_weight [cout][cin][k][k], k in {1,3,7};
 * d_out [n][H][W][round_up(cout,4)]) through conv engine `mode` (0 CUDA cores, 1 tcgen05); the launch is repeated `repeat`
 * times after one warm-up and the mean CUDA-event time per launch (ms) is written to h_ms (may be NULL) */
int fvp_debug_conv(fvp_ctx* ctx, const float* d_in, int n, int H, int W, int cin, const float* h_weight, const float* h_bias,
                   int cout, int k, int relu, int mode, float* d_out, int repeat, float* h_ms, uintptr_t stream);
/* test hook, host only (no context, no CUDA call): the launch plan of the tensor-core convolution for one layer of
 * n images H x W, cin (+ cin2 of a fused 1x1 skip conv) -> cout channels, k in {1,3,7}; engine 2 = fp16 hi/lo split,
 * 1 = 3xTF32.  out[10] = N-tile width, N tiles, weight residency (0 streamed / 1 whole image / 2 one N tile per CTA),
 * A stages, B stages, dynamic shared memory, two CTAs per SM (0/1), grid, work items, TMEM columns; out[0] = 0 when the
 * layer is left to the CUDA-core kernel.  Returns 0, or -1 on invalid arguments. */
int fvp_debug_conv_plan(int n, int H, int W, int cin, int cin2, int cout, int k, int engine, int num_sms, int* out);
/* test hook, host only: the fp16 hi/lo weight image of the tensor-core convolution (csrc/fvp_params.cu pack_tc16) for a
 * BN-folded GEMM matrix w_rows [(k*k*round_up(cin,16) + round_up(cin2,16))][coutp]; variant 0/1/2 = N tiles of up to
 * 128/32/64 columns, cb = 32 or 16 channels per K-block.  Layout contract (what k_conv_tc's descriptors read): blocks in the
 * order [phase][K-block][tap][N tile][hi, lo]; a block = n_tile rows of cb halves, the 16-byte chunks of row n XOR-ed with
 * (n >> 1) & 3 (cb = 32, SWIZZLE_64B) or (n >> 2) & 1 (cb = 16, SWIZZLE_32B); hi = fp16(w), lo = fp16((w - hi) * 2^11).
 * Returns 0, -1 on bad arguments, -2 (with *n_halves set) when capacity is too small. */
int fvp_debug_pack_tc16(const float* w_rows, int cin, int cin2, int coutp, int k, int variant, int cb, unsigned short* out,
                        long long capacity, long long* n_halves);
/* K0: [batch][V][J][H][W] -> internal channel-last, zero-bordered copy */
int fvp_stage_heatmaps(fvp_ctx* ctx, const float* d_heatmaps, int batch, uintptr_t stream);
/* K1: ProjectLayer(whole).forward + CenterNet's z-max (project_whole.py:62-88, cnns_2d.py:174)
 * -> d_plane [batch][J][X][Y] (reference layout).  Requires fvp_stage_heatmaps. */
int fvp_hdn_project(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, float* d_plane, uintptr_t stream);
/* CenterNet trunk + heads on the internal plane (cnns_2d.py:173-178)
 * d_plane_in may be NULL (use K1's result) or [batch][J][X][Y]; outputs hm [batch][X][Y], size [batch][2][X][Y] */
int fvp_center_net(fvp_ctx* ctx, const float* d_plane_in, int batch, float* d_hm, float* d_size, uintptr_t stream);
/* nms2D + top-k (core/proposal.py:13-33): d_hm [batch][X][Y] -> conf [batch][P], flat index [batch][P] (int32) */
int fvp_nms_topk(fvp_ctx* ctx, const float* d_hm, int batch, float* d_conf2d, int32_t* d_flat, uintptr_t stream);
/* K2 + C2CNet + ProposalLayer (human_detection_net.py:88-102): given top-k (conf, flat) and the size
 * map -> z-columns [batch*P][J][Z], 1-D heatmaps [batch*P][Z], proposal_centers [batch][P][7] */
int fvp_proposals(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, const float* d_conf2d,
                  const int32_t* d_flat, const float* d_size, float* d_cols, float* d_hm1d,
                  float* d_centers, uintptr_t stream);
/* K3: ProjectLayer(individual).forward + three-plane max (project_individual.py:96-136,
 * joint_localization_net.py:80-81): d_centers [batch][P][7] -> d_planes [3][batch*P][J][64][64],
 * d_offset [batch*P][3]; invalid slots (flag < 0) give zero planes. */
int fvp_jln_project(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, const float* d_centers,
                    float* d_planes, float* d_offset, uintptr_t stream);
/* P2PNet (cnns_2d.py:131-135) on planes [n][J][64][64] (NULL = K3's internal result) -> [n][J][64][64];
 * d_valid [n] int32 (NULL = all) selects images to compute */
int fvp_p2p_net(fvp_ctx* ctx, const float* d_planes, int n, const int32_t* d_valid, float* d_feat, uintptr_t stream);
/* SoftArgmaxLayer + WeightNet + fuse_pose_preds (joint_localization_net.py:20-62, weight_net.py:69-80):
 * d_feat [3][n][J][64][64], d_offset [n][3] -> pose [3][n][J][2] (offset added), conf [n],
 * weights [3][n][J], fused [n][J][3] */
int fvp_pose_head(fvp_ctx* ctx, const float* d_feat, const float* d_offset, int n, float* d_pose,
                  float* d_conf, float* d_weights, float* d_fused, uintptr_t stream);
/* C2CNet alone (cnns_1d.py:128-132): [n][J][Z] -> [n][Z] */
int fvp_c2c_net(fvp_ctx* ctx, const float* d_cols, int n, float* d_hm1d, uintptr_t stream);

/* ---- N1: heat-map renderer, the step before the hot path when TEST_HEATMAP_SRC is 'pred' or 'gt' --------
 * replaces JointsDataset.generate_input_heatmap + compute_human_scale (lib/dataset/JointsDataset.py:271-337,197-203,
 * evaluation branch) called from __getitem__ (:144-190), once per view.
 * h_joints     [batch][V][max_people][J][2] float64, joint positions in IMAGE_SIZE pixels (i.e. after the resize affine of
 *              JointsDataset.py:150 / :178, which stays on the host)
 * h_num_people [batch][V] people present in every view (0..max_people); a view without people gives a zero map
 * h_vis        [batch][V][max_people][J] 0/1 visibility (the 'gt' source) or NULL = all drawn (the 'pred' source)
 * sigma        NETWORK.SIGMA
 * d_heatmaps   [batch][V][J][H][W] fp32 - the input_heatmaps layout fvp_forward consumes.
 * Arithmetic: person scale, patch bounds and the Gaussian in IEEE float64 as NumPy 2 evaluates the reference, rounded once
 * to float32 (see oracle/heatmap_oracle.py). */
int fvp_render_heatmaps(fvp_ctx* ctx, const double* h_joints, const int32_t* h_num_people, const uint8_t* h_vis, int batch,
                        int max_people, double sigma, float* d_heatmaps, uintptr_t stream);

/* ---- N2: PoseResNet backbone ------------------------------------------------------------------------------
 * The step in front of the hot path when TEST_HEATMAP_SRC = 'image': faster_voxelpose.py:36-38 calls backbone(views[:, c]) per
 * camera; the backbone is models.resnet.get(cfg) (lib/models/resnet.py:98-215, run/validate.py:69-74).  The object takes the
 * reference's backbone state_dict (338 keys for ResNet-50) as it is.  Supported: NUM_LAYERS 18/34/50/101/152, three 4x4
 * transposed convolutions, 1x1 final layer (the reference's configs); image sides multiples of 32. */
typedef struct fvp_backbone fvp_backbone;
/* replaces models.resnet.get(cfg) (resnet.py:211-215) */
int fvp_backbone_create(int num_layers, int num_joints, int max_images, int max_h, int max_w, int device, fvp_backbone** out);
void fvp_backbone_destroy(fvp_backbone* bb);
const char* fvp_backbone_last_error(const fvp_backbone* bb);   /* bb may be NULL: error of the last failed create */
/* replaces load_state_dict (run/validate.py:72): one fp32 tensor of the reference's state_dict, reference shape */
int fvp_backbone_set_param(fvp_backbone* bb, const char* name, const float* h_data, int64_t numel);
int fvp_backbone_finalize(fvp_backbone* bb);                   /* fold BatchNorm, pack for the tensor-core engine, upload */
/* replaces ResNet.forward (resnet.py:188-201): d_images [n][3][h][w] fp32 (normalised as the reference's transform does)
 * -> d_heatmaps [n][J][h/4][w/4] fp32 */
int fvp_backbone_forward(fvp_backbone* bb, const float* d_images, int n, int h, int w, float* d_heatmaps, uintptr_t stream);
/* taps for the parity tests: number of stages, and the NCHW fp32 output of one stage - 0 = after the max-pool
 * [n][64][h/4][w/4]; 1..B = after residual block stage-1 in network order (B blocks in all); B+1..B+3 = after a transposed
 * convolution (+BN+ReLU); B+4 = the heat maps (= fvp_backbone_forward).  Valid after fvp_backbone_finalize. */
int fvp_backbone_num_stages(const fvp_backbone* bb);
int fvp_backbone_forward_slice(fvp_backbone* bb, const float* d_images, int n, int h, int w, int stage, float* d_out, uintptr_t stream);

/* convolution engine of CenterNet / P2PNet: 0 = exact-fp32 CUDA-core implicit GEMM, 1 = tcgen05/TMEM implicit GEMM with
 * the error-compensated 3xTF32 split (default set at build time) */
int fvp_set_conv_mode(fvp_ctx* ctx, int mode);

/* Latency or throughput kernels where the two differ (today: the proposal stage, lib/models/human_detection_net.py:88-102 -
 * C2CNet on one 8-CTA cluster per column, 2.4x shorter at batch 1, or on one CTA per column, half the SM time).
 * mode 1 = latency, 0 = throughput, -1 = automatic (default): latency while the context runs alone, throughput once lanes
 * (fvp_create_lane) share its device.  Results are identical within the fp32 reassociation of the C2CNet sums. */
int fvp_set_latency_mode(fvp_ctx* ctx, int mode);

/* Range guard of the default convolution engine (fp16 hi/lo split: operands must stay below 65504).  BN-folded weights
 * are checked when they are packed (a layer outside the range silently runs on the 3xTF32 engine); activations are checked
 * by the kernels as they are stored.  fvp_forward_host / fvp_wait report a violation as FVP_E_RANGE; after the
 * stream-ordered fvp_forward call fvp_check_range once the stream was synchronised (it does not synchronise itself).
 * Returns FVP_OK or FVP_E_RANGE (and clears the flag, which is shared by all contexts of a device). */
int fvp_check_range(fvp_ctx* ctx);
/* number of CenterNet / P2PNet layers whose BN-folded weights left the fp16 range at the last fvp_finalize_params and
 * therefore run on the 3xTF32 engine (0 for well-conditioned checkpoints) */
int fvp_fp16_fallback_layers(const fvp_ctx* ctx);

/* ---- introspection ---------------------------------------------------------------------------- */
/* number of kernel launches the last fvp_forward enqueued (graph replay counts its kernel nodes) */
int fvp_last_launch_count(const fvp_ctx* ctx);
/* CUDA-event time (ms) of named stages of the last forward run with fvp_set_profiling(ctx,1):
 * 0 stage-heatmaps, 1 hdn-project, 2 center-net, 3 nms, 4 proposals(c2c), 5 jln-project, 6 p2p-net,
 * 7 pose-head, 8 total */
int fvp_set_profiling(fvp_ctx* ctx, int enable);
int fvp_stage_times_ms(const fvp_ctx* ctx, float out[9]);
/* algorithmic HBM bytes per frame of K0+K1 and K3 for the current config */
int fvp_algorithmic_bytes(const fvp_ctx* ctx, int num_valid_people, double* k1_bytes, double* k3_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FVP_B200_H_ */
