// fvp_api.cu - the C ABI of libfvp_b200.so (include/fvp_b200.h): context, geometry, calibration,
// the forward pipeline (optionally replayed as a CUDA graph) and the per-stage entry points.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "fvp_ctx.h"

#include <mutex>

static std::string g_create_error;

// Range-guard status word of every device (mapped pinned host memory the conv kernels write through a __device__ pointer):
// allocated when the first context of a device is created, shared by all contexts of that device, never freed.
static std::mutex g_status_mutex;
static int* g_status_host[64] = {nullptr};

int fvp_fail(fvp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}
int fvp_fail_cuda(fvp_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
  return fvp_fail(ctx, FVP_E_CUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
}

namespace {

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
  return cudaMalloc((void**)p, count * sizeof(T));
}

// ATen's scalar linspace (the CUDA kernel the reference runs on cfg.DEVICE): symmetric about the middle
void linspace_plus(float start, float end, int n, float center, float* out) {
  const float step = n > 1 ? (end - start) / (float)(n - 1) : 0.f;
  const int half = n / 2;
  for (int i = 0; i < n; ++i) {
    volatile float v = i < half ? start + step * (float)i : end - step * (float)(n - i - 1);
    volatile float w = v + center;
    out[i] = w;
  }
}

fvp_ctx* sync_shared(fvp_ctx* ctx);
int check_stage_ready(fvp_ctx* ctx, int batch) {
  if (!ctx) return FVP_E_INVALID;
  sync_shared(ctx);                               // a lane picks up its root's current weights / tables
  if (batch < 1 || batch > ctx->cfg.max_batch)
    return fvp_fail(ctx, FVP_E_INVALID, "batch %d outside [1, max_batch=%d]", batch, ctx->cfg.max_batch);
  return FVP_OK;
}

int refuse_if_tickets_open(fvp_ctx* ctx);
// stage entry points share d_hm_cl and the workspaces with in-flight fvp_submit_host tickets: refuse before any enqueue
int check_stage_entry(fvp_ctx* ctx, int batch) {
  const int rc = check_stage_ready(ctx, batch);
  return rc != FVP_OK ? rc : refuse_if_tickets_open(ctx);
}

int upload_frame_seq(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, cudaStream_t st) {
  bool same = true;
  std::vector<int> want(batch);
  const fvp_ctx* owner = ctx->root ? ctx->root : ctx;        // calibrations and sample grids live in the root context
  for (int b = 0; b < batch; ++b) {
    const int s = h_seq_slots ? h_seq_slots[b] : 0;
    if (s < 0 || s >= owner->cfg.max_sequences || !owner->seq_set[s] || !owner->grid_ready[s])
      return fvp_fail(ctx, FVP_E_CALIB, "missing camera parameters for the current sequence (slot %d of frame %d)", s, b);
    if (ctx->h_frame_seq[b] != s || b >= ctx->frame_seq_uploaded) same = false;
    want[b] = s;
  }
  if (same) return FVP_OK;                       // device copy already holds these slots
  FVP_CUDA_OK(cudaStreamSynchronize(st));        // the pinned staging buffer may still be in flight
  for (int b = 0; b < batch; ++b) ctx->h_frame_seq[b] = want[b];
  ctx->frame_seq_uploaded = batch;
  FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_frame_seq, ctx->h_frame_seq, batch * sizeof(int), cudaMemcpyHostToDevice, st));
  return FVP_OK;
}

// depth parts per person for K3: enough CTAs for >= 6 waves of 4 CTAs / SM at small batch, so that the last, partly filled
// wave costs a few per cent (batch 1: 8 parts = 5120 CTAs = 8.6 waves, 0.124 ms; with 4 parts = 4.3 waves it was 0.134 ms)
int fvp_k3_parts(const fvp_ctx* ctx, int batch) {
  const int base = fvp_k3_patches(ctx->geom.JG) * batch * ctx->geom.P;
  int parts = 1;
  while (parts < 8 && base * parts < 24 * ctx->num_sms) parts *= 2;
  return parts;
}

// Per-context workspaces, streams and events (everything a lane owns), sized for ctx->cfg.max_batch.
cudaError_t alloc_workspaces(fvp_ctx* ctx) {
  const fvp_config& c = ctx->cfg;
  const FvpGeom& g = ctx->geom;
  const FvpProj& P = g.proj;
  const int MB = c.max_batch, n = MB * g.P, XY = g.X * g.Y, JP = P.JP;
  cudaError_t e = cudaSuccess;
  bool ok = true;
  auto A = [&](cudaError_t r) { if (r != cudaSuccess && ok) { ok = false; e = r; } };
  A(dalloc(&ctx->d_hm_in, (size_t)MB * g.V * g.J * P.H * P.W));
  A(dalloc(&ctx->d_hm_cl, (size_t)MB * g.V * g.view_stride4 * 4));
  A(dalloc(&ctx->d_plane_cl, (size_t)MB * XY * JP));
  A(dalloc(&ctx->d_hmsize, (size_t)MB * 3 * XY));
  A(dalloc(&ctx->d_conf2d, (size_t)n));
  A(dalloc(&ctx->d_flat, (size_t)n));
  A(dalloc(&ctx->d_centers, (size_t)n * 7));
  A(dalloc(&ctx->d_people, (size_t)n));
  A(dalloc(&ctx->d_img_valid, (size_t)3 * n));
  A(dalloc(&ctx->d_planes_cl, (size_t)3 * n * 4096 * JP));
  A(dalloc(&ctx->d_feat, (size_t)3 * n * g.J * 4096));
  A(dalloc(&ctx->d_pose, (size_t)3 * n * g.J * 2));
  A(dalloc(&ctx->d_maxw, (size_t)3 * n * g.J));
  A(dalloc(&ctx->d_wts, (size_t)3 * n * g.J));
  A(dalloc(&ctx->d_fused, (size_t)n * g.J * 3));
  A(dalloc(&ctx->d_conf, (size_t)n));
  A(dalloc(&ctx->d_out_fused, (size_t)n * g.J * 5));
  A(dalloc(&ctx->d_out_plane, (size_t)3 * n * g.J * 2));
  A(dalloc(&ctx->d_out_centers, (size_t)n * 7));
  ctx->tmp_floats = (size_t)3 * n * 4096 * JP;
  if ((size_t)MB * XY * JP > ctx->tmp_floats) ctx->tmp_floats = (size_t)MB * XY * JP;
  A(dalloc(&ctx->d_tmp, ctx->tmp_floats));
  for (int i = 0; i < 6; ++i) {
    A(dalloc(&ctx->cn_buf[i], (size_t)MB * XY * 64));
    A(dalloc(&ctx->p2p_buf[i], (size_t)3 * n * 4096 * 64));
  }
  A(dalloc(&ctx->d_frame_seq, (size_t)MB));
  A(cudaMallocHost((void**)&ctx->h_frame_seq, MB * sizeof(int)));
  for (int i = 0; i < 10; ++i) A(cudaEventCreate(&ctx->ev[i]));
  A(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  A(dalloc(&ctx->d_hm_in_b, (size_t)MB * g.V * g.J * P.H * P.W));
  for (int i = 0; i < 2; ++i) {
    A(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&ctx->ev_k0[i], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
  }
  if (ok) {
    A(cudaMemset(ctx->d_hm_cl, 0, (size_t)MB * g.V * g.view_stride4 * 16));   // zero borders (never written again)
    A(cudaMemset(ctx->d_people, 0, (size_t)n * sizeof(FvpPerson)));
    A(cudaMemset(ctx->d_img_valid, 0, (size_t)3 * n * sizeof(int)));
    A(cudaMemset(ctx->d_feat, 0, (size_t)3 * n * g.J * 4096 * sizeof(float)));
  }
  return e;
}

// A lane mirrors the read-only tables of its root; the root bumps shared_gen whenever weights, axes or calibrations change.
// Called at the start of every entry point that launches kernels.  Returns the context that owns the shared state.
fvp_ctx* sync_shared(fvp_ctx* ctx) {
  fvp_ctx* r = ctx->root;
  if (!r) return ctx;
  if (ctx->shared_gen != r->shared_gen) {
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }   // may hold stale weight pointers
    ctx->w_center = r->w_center; ctx->w_p2p = r->w_p2p; ctx->w_c2c = r->w_c2c; ctx->w_pose = r->w_pose;
    ctx->params_ready = r->params_ready;
    ctx->geom = r->geom;
    ctx->prop = r->prop;
    ctx->frame_seq_uploaded = 0;
    ctx->shared_gen = r->shared_gen;
  }
  return r;
}

int refuse_on_lane(fvp_ctx* ctx, const char* what) {
  if (ctx->root) return fvp_fail(ctx, FVP_E_STATE, "%s must be called on the root context, not on a lane", what);
  return FVP_OK;
}

// per-frame latency before SM efficiency?  Automatic: yes while the context runs alone, no once it shares the device with lanes
static int prefer_latency(const fvp_ctx* ctx) {
  if (ctx->latency_mode >= 0) return ctx->latency_mode;
  return ctx->root == nullptr && ctx->lanes_alive == 0;
}

// a captured graph holds the kernel choice: drop it when the automatic choice may have changed
static void drop_auto_graph(fvp_ctx* ctx) {
  if (ctx->latency_mode < 0 && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
}

FvpLaunchEnv launch_env(fvp_ctx* ctx) { return FvpLaunchEnv{ctx->num_sms, ctx->conv_mode, nullptr, nullptr, ctx->split_activations, &ctx->launch_error}; }

// Other entry points must not touch the shared workspaces while fvp_submit_host tickets are in flight.
int refuse_if_tickets_open(fvp_ctx* ctx) {
  if (ctx->tickets != ctx->waited)
    return fvp_fail(ctx, FVP_E_STATE, "fvp_submit_host tickets are outstanding: call fvp_wait before other entry points");
  return FVP_OK;
}

// After a synchronise: did any convolution of this device store an activation outside the fp16 range of the hi/lo engine?
int check_range_status(fvp_ctx* ctx) {
  if (ctx->h_status && *(volatile int*)ctx->h_status) {
    *(volatile int*)ctx->h_status = 0;
    if (ctx->conv_mode != 2) return FVP_OK;       // no layer of this context uses fp16 operands: the mark is irrelevant
    return fvp_fail(ctx, FVP_E_RANGE, "a convolution produced an activation outside the fp16 range (|x| >= 65504 or NaN): the "
                    "fp16 hi/lo tensor-core engine cannot represent it; use fvp_set_conv_mode(ctx, 1) (3xTF32) or 0 (fp32)");
  }
  return FVP_OK;
}

FvpPropArgs prop_args(fvp_ctx* ctx) {
  FvpPropArgs a = ctx->prop;
  a.g = ctx->geom;
  a.hm_cl = ctx->d_hm_cl;
  a.frame_seq = ctx->d_frame_seq;
  return a;
}

struct StageTimer {
  fvp_ctx* ctx;
  cudaStream_t st;
  void mark(int i) {
    if (ctx->profiling) cudaEventRecord(ctx->ev[i], st);
  }
};

// K1 .. finalize on internal buffers.  Outputs go to the given device pointers (may be NULL).
int run_pipeline(fvp_ctx* ctx, int batch, float* d_fused_poses, float* d_plane_poses, float* d_centers_out,
                 cudaStream_t st, int* launches) {
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P, XY = g.X * g.Y;
  StageTimer T{ctx, st};
  T.mark(1);
  fvp_launch_hdn_project(g, ctx->d_hm_cl, ctx->d_frame_seq, ctx->d_plane_cl, batch, st); ++*launches;
  T.mark(2);
  fvp_run_trunk2d(ctx->w_center, ctx->d_plane_cl, g.proj.JP, batch, g.X, g.Y, ctx->cn_buf, nullptr, true,
                  ctx->d_hmsize, 3, launches, st, launch_env(ctx));
  T.mark(3);
  fvp_launch_nms_topk(ctx->d_hmsize, (size_t)3 * XY, g.X, g.Y, g.P, batch, ctx->d_conf2d, ctx->d_flat, st); ++*launches;
  T.mark(4);
  {
    FvpPropArgs a = prop_args(ctx);
    a.conf2d = ctx->d_conf2d;
    a.flat = ctx->d_flat;
    a.size = ctx->d_hmsize + XY;
    a.size_img_stride = (size_t)3 * XY;
    a.cols_in = nullptr; a.cols_out = nullptr; a.hm1d_out = nullptr;
    a.centers = ctx->d_centers;
    a.people = ctx->d_people;
    a.img_valid = ctx->d_img_valid;
    a.n_slots = n;
    a.mode = 0;
    fvp_launch_proposals(a, ctx->w_c2c, n, prefer_latency(ctx), st); ++*launches;
  }
  T.mark(5);
  fvp_launch_jln_project(g, ctx->d_hm_cl, ctx->d_people, ctx->d_planes_cl, batch, fvp_k3_parts(ctx, batch), st);
  ++*launches;                                   // (+ one memset node)
  T.mark(6);
  fvp_run_trunk2d(ctx->w_p2p, ctx->d_planes_cl, g.proj.JP, 3 * n, 64, 64, ctx->p2p_buf, ctx->d_img_valid, false,
                  ctx->d_feat, g.J, launches, st, launch_env(ctx));
  T.mark(7);
  fvp_launch_pose_head(g, ctx->w_pose, ctx->d_feat, ctx->d_people, nullptr, n, ctx->cfg.beta, ctx->d_pose,
                       ctx->d_maxw, ctx->d_wts, ctx->d_fused, st); *launches += 2;   // k_pose_head + k_fuse
  fvp_launch_finalize(g, ctx->d_people, ctx->d_maxw, ctx->d_pose, ctx->d_fused, ctx->d_centers, batch, ctx->d_conf,
                      d_fused_poses, d_plane_poses, d_centers_out, st); ++*launches;
  T.mark(8);
  if (ctx->launch_error) {
    ctx->launch_error = 0;
    return fvp_fail(ctx, FVP_E_CUDA, "a convolution launch could not be prepared (TMA descriptor encode failed; see stderr)");
  }
  return FVP_OK;
}

void collect_times(fvp_ctx* ctx, cudaStream_t st) {
  if (!ctx->profiling) return;
  cudaStreamSynchronize(st);
  for (int i = 0; i < 8; ++i) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
  cudaEventElapsedTime(&ctx->stage_ms[8], ctx->ev[0], ctx->ev[8]);
}

}  // namespace

extern "C" {

int fvp_abi_version(void) { return FVP_ABI_VERSION; }

const char* fvp_last_error(const fvp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fvp_create(const fvp_config* cfg, int device, fvp_ctx** out) {
  if (!cfg || !out) return fvp_fail(nullptr, FVP_E_INVALID, "null argument");
  *out = nullptr;
  const fvp_config& c = *cfg;
  if (c.num_views < 1 || c.num_views > FVP_MAX_VIEWS) return fvp_fail(nullptr, FVP_E_INVALID, "num_views must be 1..%d", FVP_MAX_VIEWS);
  if (c.num_joints < 1 || c.num_joints > 20) return fvp_fail(nullptr, FVP_E_INVALID, "num_joints must be 1..20");
  if (c.max_people < 1 || c.max_people > FVP_MAX_PEOPLE) return fvp_fail(nullptr, FVP_E_INVALID, "max_people must be 1..%d", FVP_MAX_PEOPLE);
  if (c.voxels[0] != c.voxels[1])   // get_index2D divides by X (core/proposal.py:16-17): only X==Y is self-consistent
    return fvp_fail(nullptr, FVP_E_INVALID, "VOXELS_PER_AXIS[0] must equal [1] (reference get_index2D quirk)");
  if (c.voxels[0] % 8 || c.voxels[2] % 4 || c.voxels[2] > 40 || c.voxels[0] > 256)
    return fvp_fail(nullptr, FVP_E_INVALID, "coarse grid must be X=Y multiple of 8 (<=256), Z multiple of 4 (<=40)");
  if (c.ind_voxels[0] != 64 || c.ind_voxels[1] != 64 || c.ind_voxels[2] != 64)
    return fvp_fail(nullptr, FVP_E_INVALID, "INDIVIDUAL_SPEC.VOXELS_PER_AXIS must be 64^3");
  if (c.feat_channels != 32 || c.hidden_channels < 1 || c.hidden_channels > 128)
    return fvp_fail(nullptr, FVP_E_INVALID, "NUM_CHANNEL_JOINT_FEAT must be 32, hidden 1..128");
  if (c.max_batch < 1 || c.max_sequences < 1) return fvp_fail(nullptr, FVP_E_INVALID, "max_batch / max_sequences must be >= 1");
  if (c.hm_w < 8 || c.hm_h < 8) return fvp_fail(nullptr, FVP_E_INVALID, "heat map too small");
  {  // the projection kernels index the staged heat maps with 32-bit byte offsets
    const double per_view = (c.hm_w * 1.1 + 6.0) * (c.hm_h * 1.1 + 6.0) * ((c.num_joints + 3) / 4);
    if (per_view * c.num_views * c.max_batch > 2.6e8)       // x 16 B < 4 GiB
      return fvp_fail(nullptr, FVP_E_INVALID, "max_batch x views x heat-map size exceeds the 32-bit tap index range");
  }

  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fvp_fail(nullptr, FVP_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fvp_fail(nullptr, FVP_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fvp_fail(nullptr, FVP_E_CUDA, "libfvp_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);

  if (device < 0 || device >= 64) return fvp_fail(nullptr, FVP_E_INVALID, "device index %d outside [0, 64)", device);
  // per-DEVICE kernel setup (function attributes are per device; repeated for every context, which is harmless)
  e = fvp_conv_tc_init_device();
  if (e == cudaSuccess) e = fvp_conv_init_device();
  if (e != cudaSuccess) return fvp_fail(nullptr, FVP_E_CUDA, "kernel attribute setup on device %d: %s", device, cudaGetErrorString(e));
  e = fvp_proposal_init_device(c.voxels[0], c.voxels[1]);
  if (e == cudaErrorInvalidConfiguration)
    return fvp_fail(nullptr, FVP_E_INVALID, "VOXELS_PER_AXIS %dx%d: the NMS / top-k kernel stages the whole X*Y map (%d B) in shared "
                    "memory, limit 227 KB per CTA", c.voxels[0], c.voxels[1], c.voxels[0] * c.voxels[1] * 4);
  if (e != cudaSuccess) return fvp_fail(nullptr, FVP_E_CUDA, "kernel attribute setup on device %d: %s", device, cudaGetErrorString(e));
  int* h_status = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_status_mutex);
    if (!g_status_host[device]) {
      int* h = nullptr;
      int* d = nullptr;
      e = cudaHostAlloc((void**)&h, sizeof(int), cudaHostAllocMapped);
      if (e == cudaSuccess) { *h = 0; e = cudaHostGetDevicePointer((void**)&d, h, 0); }
      if (e == cudaSuccess) e = fvp_conv_tc_set_status_ptr(d);
      if (e == cudaSuccess) e = fvp_conv_set_status_ptr(d);
      if (e != cudaSuccess) return fvp_fail(nullptr, FVP_E_CUDA, "range-guard status word on device %d: %s", device, cudaGetErrorString(e));
      g_status_host[device] = h;
    }
    h_status = g_status_host[device];
  }

  fvp_ctx* ctx = new fvp_ctx();
  ctx->cfg = c;
  ctx->device = device;
  ctx->h_status = h_status;
  if (const char* sa = std::getenv("FVP_SPLIT_ACT")) ctx->split_activations = std::atoi(sa) != 0;   // A/B switch (tools only)
  fvp_build_param_table(ctx);

  FvpGeom& g = ctx->geom;
  FvpProj& P = g.proj;
  g.V = c.num_views; g.J = c.num_joints;
  P.JP = fvp_round_up(c.num_joints, 4);
  g.JG = P.JP / 4;
  g.X = c.voxels[0]; g.Y = c.voxels[1]; g.Z = c.voxels[2];
  g.P = c.max_people;
  P.W = c.hm_w; P.H = c.hm_h;
  P.PADX = (int)std::ceil(0.05 * (c.hm_w - 1)) + 2;
  P.PADY = (int)std::ceil(0.05 * (c.hm_h - 1)) + 2;
  P.WP = P.W + 2 * P.PADX; P.HP = P.H + 2 * P.PADY;
  P.ori_max = c.ori_w > c.ori_h ? c.ori_w : c.ori_h;
  P.hm_w = (float)c.hm_w; P.hm_h = (float)c.hm_h;
  P.img_w = c.image_w; P.img_h = c.image_h;
  P.wm1 = (float)(c.hm_w - 1); P.hm1 = (float)(c.hm_h - 1);
  P.r_img_w = (float)(1.0 / (double)P.img_w); P.r_img_h = (float)(1.0 / (double)P.img_h);
  P.r_wm1 = (float)(1.0 / (double)P.wm1); P.r_hm1 = (float)(1.0 / (double)P.hm1);
  g.view_stride4 = (size_t)P.HP * P.WP * g.JG;

  FvpPropArgs& pa = ctx->prop;
  memset(&pa, 0, sizeof(pa));
  pa.min_score = c.min_score;
  for (int d = 0; d < 3; ++d) {
    // fine = int(whole/ind*(vox-1)) + 1 ; scale = (fine-1)/whole ; bias = -ind/2/whole*(fine-1) - scale*(center - whole/2)
    volatile float r = c.space_size[d] / c.ind_space_size[d];
    r = r * (float)(c.ind_voxels[d] - 1);
    g.fine[d] = (int)r + 1;
    volatile float sc = (float)(g.fine[d] - 1) / c.space_size[d];
    volatile float a = -c.ind_space_size[d];
    a = a / 2.0f; a = a / c.space_size[d]; a = a * (float)(g.fine[d] - 1);
    volatile float h = c.space_size[d] / 2.0f;
    volatile float bb = c.space_center[d] - h;
    bb = sc * bb;
    volatile float bias = a - bb;
    pa.jln_scale[d] = sc; pa.jln_bias[d] = bias;
    volatile float hs = c.space_size[d] / (float)(c.voxels[d] - 1);
    volatile float hb = c.space_center[d] - h;
    pa.hdn_scale[d] = hs; pa.hdn_bias[d] = hb;
    pa.whole[d] = c.space_size[d]; pa.ind[d] = c.ind_space_size[d]; pa.ind_vox[d] = c.ind_voxels[d];
  }

  ctx->num_sms = prop.multiProcessorCount;
  // shared (read-only during forwards) state: axis tables, calibration blocks, sample-grid caches
  const size_t n_axes = (size_t)g.X + g.Y + g.Z + g.fine[0] + g.fine[1] + g.fine[2] + 192;
  bool ok = true;
  auto A = [&](cudaError_t r) { if (r != cudaSuccess && ok) { ok = false; e = r; } };
  A(dalloc(&ctx->d_axes, n_axes));
  A(dalloc(&ctx->d_coarse_grid, (size_t)c.max_sequences * g.V * g.X * g.Y * g.Z));
  A(dalloc(&ctx->d_fine_grid, (size_t)c.max_sequences * g.V * g.fine[0] * g.fine[1] * g.fine[2]));
  A(dalloc(&ctx->d_seqs, (size_t)c.max_sequences));
  if (ok) A(cudaMemset(ctx->d_seqs, 0, (size_t)c.max_sequences * sizeof(FvpSeq)));
  if (ok) A(alloc_workspaces(ctx));
  if (!ok) {
    fvp_fail(nullptr, FVP_E_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    fvp_destroy(ctx);
    return FVP_E_CUDA;
  }
  ctx->seq_set.assign(c.max_sequences, 0);
  ctx->grid_ready.assign(c.max_sequences, 0);
  g.coarse_axes = ctx->d_axes;
  g.fine_axes = ctx->d_axes + g.X + g.Y + g.Z;
  g.ind_axes = g.fine_axes + g.fine[0] + g.fine[1] + g.fine[2];
  g.coarse_grid = ctx->d_coarse_grid;
  g.fine_grid = ctx->d_fine_grid;
  g.seqs = ctx->d_seqs;
  int rc = fvp_set_axes(ctx, nullptr, nullptr, nullptr);
  if (rc != FVP_OK) {
    g_create_error = ctx->err;
    fvp_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return FVP_OK;
}

int fvp_create_lane(fvp_ctx* root, int max_batch, fvp_ctx** out) {
  if (!root || !out) return fvp_fail(root, FVP_E_INVALID, "null argument");
  *out = nullptr;
  if (root->root) return fvp_fail(root, FVP_E_STATE, "fvp_create_lane: the parent must be a root context, not a lane");
  if (max_batch < 1) return fvp_fail(root, FVP_E_INVALID, "max_batch must be >= 1");
  {
    const fvp_config& c = root->cfg;
    const double per_view = (c.hm_w * 1.1 + 6.0) * (c.hm_h * 1.1 + 6.0) * ((c.num_joints + 3) / 4);
    if (per_view * c.num_views * max_batch > 2.6e8)
      return fvp_fail(root, FVP_E_INVALID, "max_batch x views x heat-map size exceeds the 32-bit tap index range");
  }
  cudaError_t e = cudaSetDevice(root->device);
  if (e != cudaSuccess) return fvp_fail(root, FVP_E_CUDA, "cudaSetDevice(%d): %s", root->device, cudaGetErrorString(e));
  fvp_ctx* ctx = new fvp_ctx();
  ctx->cfg = root->cfg;
  ctx->cfg.max_batch = max_batch;
  ctx->device = root->device;
  ctx->h_status = root->h_status;
  ctx->num_sms = root->num_sms;
  ctx->conv_mode = root->conv_mode;
  ctx->split_activations = root->split_activations;
  ctx->root = root;
  ctx->shared_gen = -1;                            // mirrors nothing yet: the first forward copies the root's tables
  ctx->geom = root->geom;
  ctx->prop = root->prop;
  e = alloc_workspaces(ctx);
  if (e != cudaSuccess) {
    fvp_fail(root, FVP_E_CUDA, "lane allocation failed: %s", cudaGetErrorString(e));
    fvp_destroy(ctx);
    return FVP_E_CUDA;
  }
  ++root->lanes_alive;
  drop_auto_graph(root);                           // the automatic latency mode of the root has just changed
  *out = ctx;
  return FVP_OK;
}

void fvp_destroy(fvp_ctx* ctx) {
  if (!ctx) return;
  if (!ctx->root && ctx->lanes_alive > 0) {      // lanes point into this context's weights / grids
    fvp_fail(ctx, FVP_E_STATE, "fvp_destroy: %d lane(s) of this context are still alive; destroy them first", ctx->lanes_alive);
    return;
  }
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  // shared state belongs to the root; a lane frees its workspaces only
  void* shared[] = {ctx->d_weights, ctx->d_c2c_plan, ctx->d_axes, ctx->d_seqs, ctx->d_coarse_grid, ctx->d_fine_grid};
  if (!ctx->root)
    for (void* p : shared)
      if (p) cudaFree(p);
  void* ptrs[] = {ctx->d_hm_in, ctx->d_hm_cl, ctx->d_plane_cl, ctx->d_hmsize,
                  ctx->d_conf2d, ctx->d_flat, ctx->d_centers, ctx->d_people, ctx->d_img_valid, ctx->d_planes_cl,
                  ctx->d_feat, ctx->d_pose, ctx->d_maxw, ctx->d_wts, ctx->d_fused, ctx->d_conf,
                  ctx->d_out_fused, ctx->d_out_plane, ctx->d_out_centers, ctx->d_tmp, ctx->d_frame_seq, ctx->d_hm_in_b,
                  ctx->d_rj, ctx->d_rn, ctx->d_rv, ctx->d_rp};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 6; ++i) {
    if (ctx->cn_buf[i]) cudaFree(ctx->cn_buf[i]);
    if (ctx->p2p_buf[i]) cudaFree(ctx->p2p_buf[i]);
  }
  if (ctx->h_frame_seq) cudaFreeHost(ctx->h_frame_seq);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
    if (ctx->ev_k0[i]) cudaEventDestroy(ctx->ev_k0[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  }
  for (int i = 0; i < 10; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->root) {
    --ctx->root->lanes_alive;
    drop_auto_graph(ctx->root);
  }
  delete ctx;
}

int fvp_param_count(const fvp_ctx* ctx) { return ctx ? (int)ctx->params.size() : 0; }
const char* fvp_param_name(const fvp_ctx* ctx, int i) {
  return (ctx && i >= 0 && i < (int)ctx->params.size()) ? ctx->params[i].name.c_str() : nullptr;
}
int64_t fvp_param_numel(const fvp_ctx* ctx, int i) {
  return (ctx && i >= 0 && i < (int)ctx->params.size()) ? ctx->params[i].numel : -1;
}

int fvp_set_param(fvp_ctx* ctx, const char* name, const float* h_data, int64_t numel) {
  if (!ctx || !name) return FVP_E_INVALID;
  if (int lrc = refuse_on_lane(ctx, "fvp_set_param")) return lrc;
  auto it = ctx->param_index.find(name);
  if (it == ctx->param_index.end()) return fvp_fail(ctx, FVP_E_NOTFOUND, "unexpected key '%s' in state_dict", name);
  FvpParam& p = ctx->params[it->second];
  if (p.is_int) return FVP_OK;
  if (numel != p.numel) return fvp_fail(ctx, FVP_E_INVALID, "size mismatch for %s: got %lld values, expected %lld", name, (long long)numel, (long long)p.numel);
  if (!h_data) return fvp_fail(ctx, FVP_E_INVALID, "null data for %s", name);
  memcpy(p.data.data(), h_data, (size_t)numel * sizeof(float));
  p.set = true;
  ctx->params_ready = false;
  return FVP_OK;
}

int fvp_finalize_params(fvp_ctx* ctx) {
  if (!ctx) return FVP_E_INVALID;
  if (int lrc = refuse_on_lane(ctx, "fvp_finalize_params")) return lrc;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ++ctx->shared_gen;                               // lanes re-mirror the weight tables (and drop their graphs)
  if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
  int rc = fvp_pack_params(ctx);
  if (rc != FVP_OK) return rc;
  ctx->w_center.front.cin = ctx->geom.proj.JP;    // activations carry JP channels (weights rows >= J are zero)
  ctx->w_p2p.front.cin = ctx->geom.proj.JP;
  return FVP_OK;
}

int fvp_fine_voxels(const fvp_ctx* ctx, int32_t out[3]) {
  if (!ctx || !out) return FVP_E_INVALID;
  for (int d = 0; d < 3; ++d) out[d] = ctx->geom.fine[d];
  return FVP_OK;
}

int fvp_set_axes(fvp_ctx* ctx, const float* h_coarse, const float* h_fine, const float* h_individual) {
  if (!ctx) return FVP_E_INVALID;
  if (int lrc = refuse_on_lane(ctx, "fvp_set_axes")) return lrc;
  cudaSetDevice(ctx->device);
  const FvpGeom& g = ctx->geom;
  const fvp_config& c = ctx->cfg;
  const int nc = g.X + g.Y + g.Z, nf = g.fine[0] + g.fine[1] + g.fine[2];
  std::vector<float> host((size_t)nc + nf + 192);
  if (h_coarse) memcpy(host.data(), h_coarse, nc * sizeof(float));
  else {
    linspace_plus(-c.space_size[0] / 2, c.space_size[0] / 2, g.X, c.space_center[0], host.data());
    linspace_plus(-c.space_size[1] / 2, c.space_size[1] / 2, g.Y, c.space_center[1], host.data() + g.X);
    linspace_plus(-c.space_size[2] / 2, c.space_size[2] / 2, g.Z, c.space_center[2], host.data() + g.X + g.Y);
  }
  float* f = host.data() + nc;
  if (h_fine) memcpy(f, h_fine, nf * sizeof(float));
  else {
    linspace_plus(-c.space_size[0] / 2, c.space_size[0] / 2, g.fine[0], c.space_center[0], f);
    linspace_plus(-c.space_size[1] / 2, c.space_size[1] / 2, g.fine[1], c.space_center[1], f + g.fine[0]);
    linspace_plus(-c.space_size[2] / 2, c.space_size[2] / 2, g.fine[2], c.space_center[2], f + g.fine[0] + g.fine[1]);
  }
  float* q = f + nf;
  if (h_individual) memcpy(q, h_individual, 192 * sizeof(float));
  else
    for (int d = 0; d < 3; ++d)
      linspace_plus(-c.ind_space_size[d] / 2, c.ind_space_size[d] / 2, 64, c.space_center[d], q + 64 * d);
  cudaDeviceSynchronize();
  FVP_CUDA_OK(cudaMemcpy(ctx->d_axes, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  // the sample grids depend on the axes: rebuild those of every populated calibration slot now (setup path, synchronous),
  // so that forwards - of this context or of its lanes - only ever read finished grids
  for (int slot = 0; slot < ctx->cfg.max_sequences; ++slot) {
    ctx->grid_ready[slot] = 0;
    if (!ctx->seq_set[slot]) continue;
    fvp_launch_build_sample_grids(ctx->geom, slot, nullptr);
    FVP_CUDA_OK(cudaGetLastError());
    ctx->grid_ready[slot] = 1;
  }
  FVP_CUDA_OK(cudaDeviceSynchronize());
  ++ctx->shared_gen;
  return FVP_OK;
}

int fvp_set_sequence(fvp_ctx* ctx, int slot, const float* h_cameras, int num_views, const float* h_resize) {
  if (!ctx || !h_cameras || !h_resize) return FVP_E_INVALID;
  if (int lrc = refuse_on_lane(ctx, "fvp_set_sequence")) return lrc;
  if (slot < 0 || slot >= ctx->cfg.max_sequences) return fvp_fail(ctx, FVP_E_NOTFOUND, "sequence slot %d outside [0,%d)", slot, ctx->cfg.max_sequences);
  if (num_views != ctx->cfg.num_views) return fvp_fail(ctx, FVP_E_CALIB, "inconsistent number of cameras (%d given, model built for %d)", num_views, ctx->cfg.num_views);
  cudaSetDevice(ctx->device);
  FvpSeq s;
  memset(&s, 0, sizeof(s));
  for (int v = 0; v < num_views; ++v) {
    const float* c = h_cameras + 21 * v;
    FvpCam& cam = s.cam[v];
    memcpy(cam.R, c, 9 * sizeof(float));
    memcpy(cam.T, c + 9, 3 * sizeof(float));
    cam.fx = c[12]; cam.fy = c[13]; cam.cx = c[14]; cam.cy = c[15];
    memcpy(cam.k, c + 16, 3 * sizeof(float));
    memcpy(cam.p, c + 19, 2 * sizeof(float));
  }
  memcpy(s.A, h_resize, 6 * sizeof(float));
  cudaDeviceSynchronize();
  FVP_CUDA_OK(cudaMemcpy(ctx->d_seqs + slot, &s, sizeof(s), cudaMemcpyHostToDevice));
  ctx->seq_set[slot] = 1;
  // sample-grid caches of this calibration (two kernels over 5 + 164 MB, once per calibration; synchronous like the upload)
  ctx->grid_ready[slot] = 0;
  fvp_launch_build_sample_grids(ctx->geom, slot, nullptr);
  FVP_CUDA_OK(cudaGetLastError());
  FVP_CUDA_OK(cudaDeviceSynchronize());
  ctx->grid_ready[slot] = 1;
  ++ctx->shared_gen;
  return FVP_OK;
}

int fvp_use_cuda_graph(fvp_ctx* ctx, int enable) {
  if (!ctx) return FVP_E_INVALID;
  ctx->use_graph = enable != 0;
  if (!enable && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
  return FVP_OK;
}

int fvp_set_conv_mode(fvp_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return FVP_E_INVALID;
  if (mode != ctx->conv_mode && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
  ctx->conv_mode = mode;
  return FVP_OK;
}

int fvp_set_latency_mode(fvp_ctx* ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return FVP_E_INVALID;
  if (mode != ctx->latency_mode && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
  ctx->latency_mode = mode;
  return FVP_OK;
}

int fvp_set_profiling(fvp_ctx* ctx, int enable) {
  if (!ctx) return FVP_E_INVALID;
  ctx->profiling = enable != 0;
  return FVP_OK;
}
int fvp_stage_times_ms(const fvp_ctx* ctx, float out[9]) {
  if (!ctx || !out) return FVP_E_INVALID;
  memcpy(out, ctx->stage_ms, sizeof(ctx->stage_ms));
  return FVP_OK;
}
int fvp_last_launch_count(const fvp_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

int fvp_algorithmic_bytes(const fvp_ctx* ctx, int nv, double* k1, double* k3) {
  if (!ctx) return FVP_E_INVALID;
  const FvpGeom& g = ctx->geom;
  const double hm = 4.0 * g.V * g.J * g.proj.H * g.proj.W;       // heat maps read once
  if (k1) *k1 = hm + 4.0 * g.J * g.X * g.Y + 84.0 * g.V;
  if (k3) *k3 = hm + nv * (12.0 * g.J * 64 * 64 + 28.0);
  return FVP_OK;
}

// ------------------------------------------------------------------------------------------------
// whole forward
// ------------------------------------------------------------------------------------------------
static int forward_device(fvp_ctx* ctx, const float* d_heatmaps, int batch, const int32_t* h_seq_slots, float* d_fused,
                          float* d_plane, float* d_centers, cudaStream_t st) {
  int rc = check_stage_ready(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  if (!d_heatmaps) return fvp_fail(ctx, FVP_E_INVALID, "null heat maps");
  if (ctx->tickets != ctx->waited && !ctx->k0_done)
    return fvp_fail(ctx, FVP_E_STATE, "fvp_submit_host tickets are outstanding: call fvp_wait before other entry points");
  cudaSetDevice(ctx->device);
  rc = upload_frame_seq(ctx, batch, h_seq_slots, st);
  if (rc != FVP_OK) return rc;
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P;
  int launches = 0;
  StageTimer T{ctx, st};
  T.mark(0);
  fvp_launch_stage_heatmaps(g, d_heatmaps, ctx->d_hm_cl, batch, st); ++launches;
  if (ctx->k0_done) FVP_CUDA_OK(cudaEventRecord(ctx->k0_done, st));   // the input buffer is free again

  // Stream capture is illegal on the legacy default stream: in graph mode run on the context's own
  // stream, ordered after / before the caller's stream with events.
  cudaStream_t caller = st;
  const bool reroute = ctx->use_graph && !ctx->profiling && (st == nullptr || st == cudaStreamLegacy);
  if (reroute) {
    FVP_CUDA_OK(cudaEventRecord(ctx->ev[9], caller));
    st = ctx->own_stream;
    FVP_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev[9], 0));
  }
  bool sig_same = ctx->graph_exec && ctx->graph_batch == batch && (int)ctx->graph_seqs.size() == batch;
  if (sig_same)
    for (int b = 0; b < batch; ++b) sig_same = sig_same && ctx->graph_seqs[b] == ctx->h_frame_seq[b];
  if (ctx->use_graph && !ctx->profiling) {
    if (!sig_same) {
      // first call with this signature: run eagerly, then capture the same sequence into a graph bound to the internal
      // output buffers.  Whatever fails inside the capture, the stream is taken OUT of capture mode again (an unfinished
      // capture would poison every later call on it, the caller's PyTorch work included) and the previous executable
      // graph is only replaced once the new one was instantiated.
      rc = run_pipeline(ctx, batch, ctx->d_out_fused, ctx->d_out_plane, ctx->d_out_centers, st, &launches);
      if (rc != FVP_OK) return rc;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      int cap_launches = 0;
      FVP_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int cap_rc = run_pipeline(ctx, batch, ctx->d_out_fused, ctx->d_out_plane, ctx->d_out_centers, st, &cap_launches);
      const cudaError_t launch_err = cudaGetLastError();
      cudaError_t cap_err = cudaStreamEndCapture(st, &graph);       // always ends the capture (graph = NULL on failure)
      if (cap_err == cudaSuccess && (cap_rc != FVP_OK || launch_err != cudaSuccess)) cap_err = launch_err != cudaSuccess ? launch_err : cudaErrorUnknown;
      if (cap_err == cudaSuccess) cap_err = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (cap_err != cudaSuccess) {
        if (exec) cudaGraphExecDestroy(exec);
        cudaGetLastError();                                         // clear the sticky-free error state of the failed capture
        return fvp_fail(ctx, FVP_E_CUDA, "CUDA graph capture of the forward failed: %s", cudaGetErrorString(cap_err));
      }
      if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
      ctx->graph_exec = exec;
      ctx->graph_batch = batch;
      ctx->graph_seqs.assign(ctx->h_frame_seq, ctx->h_frame_seq + batch);
      ctx->graph_launches = cap_launches;
    } else {
      FVP_CUDA_OK(cudaGraphLaunch(ctx->graph_exec, st));
      launches += ctx->graph_launches;
    }
    if (d_fused) FVP_CUDA_OK(cudaMemcpyAsync(d_fused, ctx->d_out_fused, (size_t)n * g.J * 5 * 4, cudaMemcpyDeviceToDevice, st));
    if (d_plane) FVP_CUDA_OK(cudaMemcpyAsync(d_plane, ctx->d_out_plane, (size_t)3 * n * g.J * 2 * 4, cudaMemcpyDeviceToDevice, st));
    if (d_centers) FVP_CUDA_OK(cudaMemcpyAsync(d_centers, ctx->d_out_centers, (size_t)n * 7 * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    rc = run_pipeline(ctx, batch, d_fused, d_plane, d_centers, st, &launches);
    if (rc != FVP_OK) return rc;
  }
  if (reroute) {
    FVP_CUDA_OK(cudaEventRecord(ctx->ev[9], st));
    FVP_CUDA_OK(cudaStreamWaitEvent(caller, ctx->ev[9], 0));
  }
  ctx->last_launches = launches;
  FVP_CUDA_OK(cudaGetLastError());
  collect_times(ctx, st);
  return FVP_OK;
}

int fvp_forward(fvp_ctx* ctx, const float* d_heatmaps, int batch, const int32_t* h_seq_slots, float* d_fused_poses,
                float* d_plane_poses, float* d_proposal_centers, uintptr_t stream) {
  if (!ctx) return FVP_E_INVALID;
  return forward_device(ctx, d_heatmaps, batch, h_seq_slots, d_fused_poses, d_plane_poses, d_proposal_centers,
                        (cudaStream_t)stream);
}

int fvp_forward_host(fvp_ctx* ctx, const float* h_heatmaps, int batch, const int32_t* h_seq_slots, float* h_fused,
                     float* h_plane, float* h_centers, uintptr_t stream) {
  if (!ctx) return FVP_E_INVALID;
  int rc = check_stage_ready(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!h_heatmaps) return fvp_fail(ctx, FVP_E_INVALID, "null heat maps");
  rc = refuse_if_tickets_open(ctx);             // before ANY enqueue: the H2D below would overwrite a ticket's input
  if (rc != FVP_OK) return rc;
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P;
  const size_t in_bytes = (size_t)batch * g.V * g.J * g.proj.H * g.proj.W * sizeof(float);
  FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_hm_in, h_heatmaps, in_bytes, cudaMemcpyHostToDevice, st));
  rc = forward_device(ctx, ctx->d_hm_in, batch, h_seq_slots, ctx->d_out_fused, ctx->d_out_plane, ctx->d_out_centers, st);
  if (rc != FVP_OK) return rc;
  if (h_fused) FVP_CUDA_OK(cudaMemcpyAsync(h_fused, ctx->d_out_fused, (size_t)n * g.J * 5 * 4, cudaMemcpyDeviceToHost, st));
  if (h_plane) FVP_CUDA_OK(cudaMemcpyAsync(h_plane, ctx->d_out_plane, (size_t)3 * n * g.J * 2 * 4, cudaMemcpyDeviceToHost, st));
  if (h_centers) FVP_CUDA_OK(cudaMemcpyAsync(h_centers, ctx->d_out_centers, (size_t)n * 7 * 4, cudaMemcpyDeviceToHost, st));
  FVP_CUDA_OK(cudaStreamSynchronize(st));
  return check_range_status(ctx);
}

// Pipelined host entry: fvp_submit_host enqueues H2D (copy stream) -> forward (context stream) -> D2H and returns a
// ticket without waiting; the copy of step i+1 overlaps the kernels of step i (two device input buffers).  Results
// land in the caller's host buffers once fvp_wait(ticket) returns.  At most two tickets may be outstanding and they
// must be waited for in order; buffers of an outstanding ticket must not be reused.  Pinned host memory is needed
// for the copies to be asynchronous (pageable memory still works, without the overlap).
int fvp_submit_host(fvp_ctx* ctx, const float* h_heatmaps, int batch, const int32_t* h_seq_slots, float* h_fused,
                    float* h_plane, float* h_centers, long long* ticket) {
  if (!ctx) return FVP_E_INVALID;
  int rc = check_stage_ready(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!h_heatmaps || !ticket) return fvp_fail(ctx, FVP_E_INVALID, "null heat maps / ticket");
  if (ctx->tickets - ctx->waited >= 2) return fvp_fail(ctx, FVP_E_STATE, "two tickets outstanding: call fvp_wait first");
  cudaSetDevice(ctx->device);
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P, p = (int)(ctx->tickets & 1);
  float* d_in = p ? ctx->d_hm_in_b : ctx->d_hm_in;
  const size_t in_bytes = (size_t)batch * g.V * g.J * g.proj.H * g.proj.W * sizeof(float);
  if (ctx->tickets >= 2) FVP_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_k0[p], 0));   // last reader of d_in
  FVP_CUDA_OK(cudaMemcpyAsync(d_in, h_heatmaps, in_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  FVP_CUDA_OK(cudaEventRecord(ctx->ev_h2d[p], ctx->copy_stream));
  cudaStream_t st = ctx->own_stream;
  FVP_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev_h2d[p], 0));
  ctx->k0_done = ctx->ev_k0[p];
  rc = forward_device(ctx, d_in, batch, h_seq_slots, ctx->d_out_fused, ctx->d_out_plane, ctx->d_out_centers, st);
  ctx->k0_done = nullptr;
  if (rc != FVP_OK) return rc;
  if (h_fused) FVP_CUDA_OK(cudaMemcpyAsync(h_fused, ctx->d_out_fused, (size_t)n * g.J * 5 * 4, cudaMemcpyDeviceToHost, st));
  if (h_plane) FVP_CUDA_OK(cudaMemcpyAsync(h_plane, ctx->d_out_plane, (size_t)3 * n * g.J * 2 * 4, cudaMemcpyDeviceToHost, st));
  if (h_centers) FVP_CUDA_OK(cudaMemcpyAsync(h_centers, ctx->d_out_centers, (size_t)n * 7 * 4, cudaMemcpyDeviceToHost, st));
  FVP_CUDA_OK(cudaEventRecord(ctx->ev_done[p], st));
  *ticket = ctx->tickets++;
  return FVP_OK;
}

int fvp_wait(fvp_ctx* ctx, long long ticket) {
  if (!ctx) return FVP_E_INVALID;
  if (ticket < ctx->waited) return FVP_OK;                               // already complete
  if (ticket >= ctx->tickets) return fvp_fail(ctx, FVP_E_INVALID, "unknown ticket %lld", ticket);
  cudaSetDevice(ctx->device);
  for (long long t = ctx->waited; t <= ticket; ++t) FVP_CUDA_OK(cudaEventSynchronize(ctx->ev_done[t & 1]));
  ctx->waited = ticket + 1;
  return check_range_status(ctx);
}

int fvp_fp16_fallback_layers(const fvp_ctx* ctx) { return ctx ? ctx->fp16_fallback_layers : 0; }

int fvp_check_range(fvp_ctx* ctx) {
  if (!ctx) return FVP_E_INVALID;
  return check_range_status(ctx);
}

// ------------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------------
int fvp_stage_heatmaps(fvp_ctx* ctx, const float* d_heatmaps, int batch, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  cudaSetDevice(ctx->device);
  fvp_launch_stage_heatmaps(ctx->geom, d_heatmaps, ctx->d_hm_cl, batch, (cudaStream_t)stream);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_debug_conv(fvp_ctx* ctx, const float* d_in, int n, int H, int W, int cin, const float* h_weight, const float* h_bias,
                   int cout, int k, int relu, int mode, float* d_out, int repeat, float* h_ms, uintptr_t stream) {
  if (!ctx || !d_in || !h_weight || !h_bias || !d_out || (k != 1 && k != 3 && k != 7) || cin % 4 || mode < 0 || (mode & 0xff) > 4) return FVP_E_INVALID;
  sync_shared(ctx);
  cudaSetDevice(ctx->device);
  return fvp_debug_conv_impl(ctx, d_in, n, H, W, cin, h_weight, h_bias, cout, k, relu, mode, d_out, repeat, h_ms, (cudaStream_t)stream);
}

int fvp_debug_project(fvp_ctx* ctx, int slot, const float* d_points, int n, float* d_ix, float* d_iy, uintptr_t stream) {
  if (!ctx || !d_points || !d_ix || !d_iy || n < 1) return FVP_E_INVALID;
  const fvp_ctx* owner = sync_shared(ctx);
  if (slot < 0 || slot >= owner->cfg.max_sequences || !owner->seq_set[slot])
    return fvp_fail(ctx, FVP_E_CALIB, "missing camera parameters for the current sequence (slot %d)", slot);
  cudaSetDevice(ctx->device);
  fvp_launch_debug_project(ctx->geom, slot, d_points, n, d_ix, d_iy, (cudaStream_t)stream);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_hdn_project(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, float* d_plane, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = upload_frame_seq(ctx, batch, h_seq_slots, st);
  if (rc != FVP_OK) return rc;
  const FvpGeom& g = ctx->geom;
  fvp_launch_hdn_project(g, ctx->d_hm_cl, ctx->d_frame_seq, ctx->d_plane_cl, batch, st);
  if (d_plane) fvp_launch_nhwc_to_nchw(ctx->d_plane_cl, d_plane, batch, g.X * g.Y, g.proj.JP, g.J, st);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_center_net(fvp_ctx* ctx, const float* d_plane_in, int batch, float* d_hm, float* d_size, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const FvpGeom& g = ctx->geom;
  const int XY = g.X * g.Y;
  if (d_plane_in) fvp_launch_nchw_to_nhwc(d_plane_in, ctx->d_plane_cl, batch, XY, g.proj.JP, g.J, st);
  int launches = 0;
  fvp_run_trunk2d(ctx->w_center, ctx->d_plane_cl, g.proj.JP, batch, g.X, g.Y, ctx->cn_buf, nullptr, true,
                  ctx->d_hmsize, 3, &launches, st, launch_env(ctx));
  if (ctx->launch_error) { ctx->launch_error = 0; return fvp_fail(ctx, FVP_E_CUDA, "a convolution launch could not be prepared (TMA descriptor)"); }
  for (int b = 0; b < batch; ++b) {
    if (d_hm) FVP_CUDA_OK(cudaMemcpyAsync(d_hm + (size_t)b * XY, ctx->d_hmsize + (size_t)b * 3 * XY, XY * 4, cudaMemcpyDeviceToDevice, st));
    if (d_size) FVP_CUDA_OK(cudaMemcpyAsync(d_size + (size_t)b * 2 * XY, ctx->d_hmsize + (size_t)b * 3 * XY + XY, 2 * XY * 4, cudaMemcpyDeviceToDevice, st));
  }
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_nms_topk(fvp_ctx* ctx, const float* d_hm, int batch, float* d_conf2d, int32_t* d_flat, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!d_hm || !d_conf2d || !d_flat) return fvp_fail(ctx, FVP_E_INVALID, "null argument");
  cudaSetDevice(ctx->device);
  const FvpGeom& g = ctx->geom;
  fvp_launch_nms_topk(d_hm, (size_t)g.X * g.Y, g.X, g.Y, g.P, batch, d_conf2d, d_flat, (cudaStream_t)stream);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_proposals(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, const float* d_conf2d, const int32_t* d_flat,
                  const float* d_size, float* d_cols, float* d_hm1d, float* d_centers, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  if (!d_conf2d || !d_flat || !d_size) return fvp_fail(ctx, FVP_E_INVALID, "null argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = upload_frame_seq(ctx, batch, h_seq_slots, st);
  if (rc != FVP_OK) return rc;
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P;
  FvpPropArgs a = prop_args(ctx);
  a.conf2d = d_conf2d; a.flat = d_flat;
  a.size = d_size; a.size_img_stride = (size_t)2 * g.X * g.Y;
  a.cols_in = nullptr; a.cols_out = d_cols; a.hm1d_out = d_hm1d;
  a.centers = d_centers ? d_centers : ctx->d_centers;
  a.people = ctx->d_people; a.img_valid = ctx->d_img_valid; a.n_slots = n;
  a.mode = 0;
  fvp_launch_proposals(a, ctx->w_c2c, n, prefer_latency(ctx), st);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_c2c_net(fvp_ctx* ctx, const float* d_cols, int n, float* d_hm1d, uintptr_t stream) {
  if (!ctx || !d_cols || !d_hm1d || n < 1) return FVP_E_INVALID;
  sync_shared(ctx);
  if (int trc = refuse_if_tickets_open(ctx)) return trc;
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  cudaSetDevice(ctx->device);
  FvpPropArgs a = prop_args(ctx);
  a.cols_in = d_cols; a.cols_out = nullptr; a.hm1d_out = d_hm1d;
  a.centers = nullptr; a.people = nullptr; a.img_valid = nullptr; a.n_slots = n;
  a.mode = 1;
  fvp_launch_proposals(a, ctx->w_c2c, n, prefer_latency(ctx), (cudaStream_t)stream);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_jln_project(fvp_ctx* ctx, int batch, const int32_t* h_seq_slots, const float* d_centers, float* d_planes,
                    float* d_offset, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!d_centers) return fvp_fail(ctx, FVP_E_INVALID, "null centers");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  rc = upload_frame_seq(ctx, batch, h_seq_slots, st);
  if (rc != FVP_OK) return rc;
  const FvpGeom& g = ctx->geom;
  const int n = batch * g.P;
  FvpPropArgs a = prop_args(ctx);
  a.people = ctx->d_people; a.img_valid = ctx->d_img_valid; a.n_slots = n;
  fvp_launch_people_from_centers(a, d_centers, n, st);
  fvp_launch_jln_project(g, ctx->d_hm_cl, ctx->d_people, ctx->d_planes_cl, batch, fvp_k3_parts(ctx, batch), st);
  if (d_planes) fvp_launch_nhwc_to_nchw(ctx->d_planes_cl, d_planes, 3 * n, 4096, g.proj.JP, g.J, st);
  if (d_offset) {
    FVP_CUDA_OK(cudaMemcpy2DAsync(d_offset, 3 * sizeof(float), (const char*)ctx->d_people + offsetof(FvpPerson, offset),
                                  sizeof(FvpPerson), 3 * sizeof(float), n, cudaMemcpyDeviceToDevice, st));
  }
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_p2p_net(fvp_ctx* ctx, const float* d_planes, int n, const int32_t* d_valid, float* d_feat, uintptr_t stream) {
  if (!ctx || n < 1 || !d_feat) return FVP_E_INVALID;
  sync_shared(ctx);
  if (int trc = refuse_if_tickets_open(ctx)) return trc;
  if (n > 3 * ctx->cfg.max_batch * ctx->geom.P) return fvp_fail(ctx, FVP_E_INVALID, "too many images (%d)", n);
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const FvpGeom& g = ctx->geom;
  const float* in = ctx->d_planes_cl;
  if (d_planes) {
    fvp_launch_nchw_to_nhwc(d_planes, ctx->d_tmp, n, 4096, g.proj.JP, g.J, st);
    in = ctx->d_tmp;
  }
  int launches = 0;
  fvp_run_trunk2d(ctx->w_p2p, in, g.proj.JP, n, 64, 64, ctx->p2p_buf, d_valid, false, d_feat, g.J, &launches, st, launch_env(ctx));
  if (ctx->launch_error) { ctx->launch_error = 0; return fvp_fail(ctx, FVP_E_CUDA, "a convolution launch could not be prepared (TMA descriptor)"); }
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

int fvp_pose_head(fvp_ctx* ctx, const float* d_feat, const float* d_offset, int n, float* d_pose, float* d_conf,
                  float* d_weights, float* d_fused, uintptr_t stream) {
  if (!ctx || !d_feat || !d_offset || n < 1) return FVP_E_INVALID;
  sync_shared(ctx);
  if (int trc = refuse_if_tickets_open(ctx)) return trc;
  if (n > ctx->cfg.max_batch * ctx->geom.P) return fvp_fail(ctx, FVP_E_INVALID, "too many persons (%d)", n);
  if (!ctx->params_ready) return fvp_fail(ctx, FVP_E_STATE, "fvp_finalize_params has not been called");
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const FvpGeom& g = ctx->geom;
  float* pose = d_pose ? d_pose : ctx->d_pose;
  float* wts = d_weights ? d_weights : ctx->d_wts;
  float* fused = d_fused ? d_fused : ctx->d_fused;
  fvp_launch_pose_head(g, ctx->w_pose, d_feat, nullptr, d_offset, n, ctx->cfg.beta, pose, ctx->d_maxw, wts, fused, st);
  if (d_conf) {
    // conf[p] = mean over (plane, joint) of the max soft-max weight: reuse the finalize kernel on a
    // temporary all-valid person table
    std::vector<FvpPerson> hp(n);
    memset(hp.data(), 0, n * sizeof(FvpPerson));
    for (int i = 0; i < n; ++i) hp[i].valid = 1;
    FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_people, hp.data(), n * sizeof(FvpPerson), cudaMemcpyHostToDevice, st));
    FVP_CUDA_OK(cudaStreamSynchronize(st));
    const int batch = fvp_cdiv(n, g.P);
    (void)batch;
    FvpGeom g1 = g;
    g1.P = 1;
    fvp_launch_finalize(g1, ctx->d_people, ctx->d_maxw, pose, fused, ctx->d_centers, n, d_conf, nullptr, nullptr, nullptr, st);
  }
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

// ------------------------------------------------------------------------------------------------
// N1: heat-map renderer (the step before the hot path for TEST_HEATMAP_SRC 'pred' / 'gt')
// ------------------------------------------------------------------------------------------------
int fvp_render_heatmaps(fvp_ctx* ctx, const double* h_joints, const int32_t* h_num_people, const uint8_t* h_vis, int batch,
                        int max_people, double sigma, float* d_heatmaps, uintptr_t stream) {
  int rc = check_stage_entry(ctx, batch);
  if (rc != FVP_OK) return rc;
  if (!h_joints || !h_num_people || !d_heatmaps) return fvp_fail(ctx, FVP_E_INVALID, "null joints / counts / output");
  if (max_people < 1 || max_people > FVP_MAX_PEOPLE)
    return fvp_fail(ctx, FVP_E_INVALID, "max_people %d outside [1, %d]", max_people, FVP_MAX_PEOPLE);
  if (!(sigma > 0.0)) return fvp_fail(ctx, FVP_E_INVALID, "sigma must be positive");
  const FvpGeom& g = ctx->geom;
  const int views = batch * g.V;
  for (int i = 0; i < views; ++i)
    if (h_num_people[i] < 0 || h_num_people[i] > max_people)
      return fvp_fail(ctx, FVP_E_INVALID, "num_people[%d] = %d outside [0, %d]", i, h_num_people[i], max_people);
  cudaSetDevice(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!ctx->d_rj) {                              // staging sized once for (max_batch, FVP_MAX_PEOPLE)
    const size_t cap = (size_t)ctx->cfg.max_batch * g.V * FVP_MAX_PEOPLE;
    FVP_CUDA_OK(cudaMalloc((void**)&ctx->d_rj, cap * g.J * 2 * sizeof(double)));
    FVP_CUDA_OK(cudaMalloc((void**)&ctx->d_rn, (size_t)ctx->cfg.max_batch * g.V * sizeof(int)));
    FVP_CUDA_OK(cudaMalloc((void**)&ctx->d_rv, cap * g.J));
    FVP_CUDA_OK(cudaMalloc(&ctx->d_rp, fvp_render_patch_bytes(ctx->cfg.max_batch * g.V, FVP_MAX_PEOPLE, g.J)));
  }
  const size_t n = (size_t)views * max_people;
  FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_rj, h_joints, n * g.J * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_rn, h_num_people, views * sizeof(int), cudaMemcpyHostToDevice, st));
  if (h_vis) FVP_CUDA_OK(cudaMemcpyAsync(ctx->d_rv, h_vis, n * g.J, cudaMemcpyHostToDevice, st));
  const double sx = (double)ctx->cfg.image_w / (double)g.proj.W, sy = (double)ctx->cfg.image_h / (double)g.proj.H;
  fvp_launch_render_heatmaps(ctx->d_rj, ctx->d_rn, h_vis ? ctx->d_rv : nullptr, views, max_people, g.J, g.proj.W, g.proj.H,
                             sx, sy, sigma, ctx->d_rp, d_heatmaps, st);
  FVP_CUDA_OK(cudaGetLastError());
  return FVP_OK;
}

}  // extern "C"
