// fvp_ctx.h - the context object behind the C ABI (host only).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "fvp_kernels.h"

struct FvpLayer {
  std::string key, bn;
  int cin, cout, k, ndim;
  bool transposed;
};
struct FvpParam {
  std::string name;
  int64_t numel;
  bool is_int;
  bool set;
  std::vector<float> data;
};

struct fvp_ctx {
  fvp_config cfg;
  int device;
  std::string err;

  // Lanes (fvp_create_lane): a lane is a context with its OWN workspaces, streams, events and CUDA graph that SHARES the
  // read-only state of its root context - packed weights, axis tables, calibration blocks, sample-grid caches.  Several
  // frames can then be in flight on one GPU (one lane each) with one copy of the weights and of the 164 MB fine grid.
  fvp_ctx* root = nullptr;            // NULL for a root context
  int lanes_alive = 0;                // root only: lanes created and not yet destroyed
  long long shared_gen = 0;           // root: bumped whenever weights / axes / calibrations change; lane: generation it mirrors

  // parameters
  std::vector<FvpLayer> layers;
  std::vector<FvpParam> params;
  std::map<std::string, int> param_index;
  bool params_ready = false;
  float* d_weights = nullptr;
  void* d_c2c_plan = nullptr;         // chunk table of the C2CNet weight ring (device)
  FvpTrunkW w_center, w_p2p;
  FvpC2CW w_c2c;
  FvpPoseW w_pose;

  // geometry
  FvpGeom geom;
  float* d_axes = nullptr;            // coarse | fine | individual
  FvpSeq* d_seqs = nullptr;
  std::vector<int> seq_set;           // which calibration slots are populated
  FvpPropArgs prop;                   // constants of the proposal kernel (pointers filled per call)

  // workspaces (sized for cfg.max_batch)
  float* d_hm_in = nullptr;           // [MB][V][J][H][W] staging for the host entry point
  float* d_hm_cl = nullptr;
  float* d_plane_cl = nullptr;        // [MB][X][Y][JP]
  float* d_hmsize = nullptr;          // [MB][3][X][Y]  (hm, size_w, size_h)
  float* d_conf2d = nullptr;          // [MB*P]
  int* d_flat = nullptr;              // [MB*P]
  float* d_centers = nullptr;         // [MB*P][7]
  FvpPerson* d_people = nullptr;      // [MB*P]
  int* d_img_valid = nullptr;         // [3*MB*P]
  float* d_planes_cl = nullptr;       // [3][MB*P][64][64][JP]
  float* d_feat = nullptr;            // [3][MB*P][J][64][64]
  float* d_pose = nullptr;            // [3][MB*P][J][2]
  float* d_maxw = nullptr;            // [3][MB*P][J]
  float* d_wts = nullptr;             // [3][MB*P][J]
  float* d_fused = nullptr;           // [MB*P][J][3]
  float* d_conf = nullptr;            // [MB*P]
  float* d_out_fused = nullptr;       // [MB][P][J][5]   (graph / host entry outputs)
  float* d_out_plane = nullptr;       // [3][MB][P][J][2]
  float* d_out_centers = nullptr;     // [MB][P][7]
  float* d_tmp = nullptr;             // scratch for stage-API layout conversions
  size_t tmp_floats = 0;
  float* cn_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* p2p_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* d_frame_seq = nullptr;         // [MB]
  int* h_frame_seq = nullptr;         // pinned
  int frame_seq_uploaded = 0;         // how many leading entries of d_frame_seq mirror h_frame_seq
  int num_sms = 148;
  int conv_mode = 2;                  // 0 = fp32 CUDA cores, 1 = tcgen05 3xTF32 (7x7 on CUDA cores), 2 = tcgen05 fp16 split
  int latency_mode = -1;              // fvp_set_latency_mode: -1 auto (on when the context has no lanes), 0 off, 1 on
  int launch_error = 0;               // set by a conv launcher that could not prepare its launch (FvpLaunchEnv::error)
  int split_activations = 1;          // engine 2: split (fp16 hi / lo) activations between layers, TMA-fed convolutions
  int fp16_fallback_layers = 0;       // conv layers packed without an fp16 image (weights outside the fp16 range)
  int* h_status = nullptr;            // range-guard word of this device (mapped pinned host memory, owned by fvp_api.cu)

  // cuda graph
  bool use_graph = false;
  cudaStream_t own_stream = nullptr;  // used for graph capture/replay when the caller passes the legacy stream
  cudaGraphExec_t graph_exec = nullptr;
  int graph_batch = 0;
  std::vector<int> graph_seqs;
  int graph_launches = 0;

  // introspection
  int last_launches = 0;
  bool profiling = false;
  cudaEvent_t ev[10] = {nullptr};
  // pipelined host entry (fvp_submit_host / fvp_wait): two input buffers, a copy stream and per-slot events
  float2* d_coarse_grid = nullptr;    // sample-grid cache, see FvpGeom
  float2* d_fine_grid = nullptr;
  std::vector<char> grid_ready;       // per slot: grids match the current cameras and axes
  float* d_hm_in_b = nullptr;         // second [MB][V][J][H][W] input buffer
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k0[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t k0_done = nullptr;      // when set, forward_device records it right after the staging kernel
  long long tickets = 0;              // submitted so far
  long long waited = 0;               // tickets already waited for
  float stage_ms[9] = {0};
  // heat-map renderer (fvp_render_heatmaps): device copies of the poses / counts / visibility and the patch table
  double* d_rj = nullptr;
  int* d_rn = nullptr;
  unsigned char* d_rv = nullptr;
  void* d_rp = nullptr;
};

void fvp_build_param_table(fvp_ctx* ctx);
int fvp_pack_params(fvp_ctx* ctx);
int fvp_debug_conv_impl(fvp_ctx* ctx, const float* d_in, int n, int H, int W, int cin, const float* h_w, const float* h_b,
                        int cout, int k, int relu, int mode, float* d_out, int repeat, float* ms_out, cudaStream_t st);
