// fvp_backbone.cu - N2: first slice of the PoseResNet backbone (lib/models/resnet.py:98-201), the step in
// front of the hot path when TEST_HEATMAP_SRC = 'image' (faster_voxelpose.py:36-38: heat maps = backbone(views[:, c])).
//
//   stem      conv 7x7 stride 2 (3 -> 64, no bias) + BatchNorm(eval) + ReLU      resnet.py:103-107,189-191
//             k_stem7x7s2: exact-fp32 CUDA-core kernel (3 input channels = 2 % of the backbone's MACs), NCHW images in,
//             NHWC out, BN folded into the weights
//   max-pool  3x3 stride 2 pad 1                                                  resnet.py:108,192
//   layer1    3 Bottleneck blocks (1x1 64->64, 3x3 64->64, 1x1 64->256, + 1x1 downsample of the block input in block 0;
//             out += residual; ReLU, resnet.py:57-95,118-146) or 2 BasicBlocks (ResNet-18/34, resnet.py:22-54)
//             on the tcgen05 engine of fvp_conv_tc.cu (fp16 hi/lo split, fp32 accumulation): BN folded, the downsample
//             conv fused as the second K segment of the block's last conv, the identity residual added in its epilogue.
// layer2-4 (stride-2 3x3 / 1x1), the 4x4 stride-2 transposed convolutions and the final 1x1 are not built yet.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "fvp_kernels.h"

extern "C" int fvp_debug_pack_tc16(const float* w_rows, int cin, int cin2, int coutp, int k, int variant, int cb, unsigned short* out,
                                   long long capacity, long long* n_halves);

namespace {

// ---- stem: one CTA = 8 x 8 output pixels x 64 channels; thread = (pixel, 16-channel group) -----------------------------
constexpr int ST_T = 8;                        // output tile edge
constexpr int ST_IN = 2 * ST_T + 5;            // input patch edge (stride 2, 7 taps)
__global__ void __launch_bounds__(256) k_stem7x7s2(const float* __restrict__ img, const float* __restrict__ w /*[147][64]*/,
                                                   const float* __restrict__ bias, float* __restrict__ out, int h, int w_in, int ho,
                                                   int wo) {
  __shared__ float s_in[3][ST_IN][ST_IN + 1];
  __shared__ __align__(16) float s_w[147 * 64];
  const int n = blockIdx.z, y0 = blockIdx.y * ST_T, x0 = blockIdx.x * ST_T, tid = threadIdx.x;
  for (int i = tid; i < 147 * 64; i += 256) s_w[i] = w[i];
  const float* im = img + (size_t)n * 3 * h * w_in;
  for (int i = tid; i < 3 * ST_IN * ST_IN; i += 256) {
    const int c = i / (ST_IN * ST_IN), r = (i / ST_IN) % ST_IN, q = i % ST_IN;
    const int gy = 2 * y0 - 3 + r, gx = 2 * x0 - 3 + q;
    s_in[c][r][q] = (gy >= 0 && gy < h && gx >= 0 && gx < w_in) ? __ldg(im + ((size_t)c * h + gy) * w_in + gx) : 0.f;
  }
  __syncthreads();
  const int p = tid & 63, cg = tid >> 6, py = p >> 3, px = p & 7;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float x = s_in[c][2 * py + dy][2 * px + dx];
        const float4* wr = reinterpret_cast<const float4*>(s_w + ((c * 7 + dy) * 7 + dx) * 64 + cg * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q] = fmaf(x, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(x, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, wv.w, acc[4 * q + 3]);
        }
      }
  const int oy = y0 + py, ox = x0 + px;
  if (oy >= ho || ox >= wo) return;
  float4* o = reinterpret_cast<float4*>(out + (((size_t)n * ho + oy) * wo + ox) * 64 + cg * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cg * 16) + q);
    o[q] = make_float4(fmaxf(acc[4 * q] + b.x, 0.f), fmaxf(acc[4 * q + 1] + b.y, 0.f), fmaxf(acc[4 * q + 2] + b.z, 0.f),
                       fmaxf(acc[4 * q + 3] + b.w, 0.f));
  }
}

// MaxPool2d(3, 2, 1) on NHWC, 4 channels per thread (padding = -inf, i.e. ignored)
__global__ void __launch_bounds__(256) k_maxpool3s2(const float4* __restrict__ in, float4* __restrict__ out, int H, int W, int Ho, int Wo,
                                                    int C4) {
  const int n = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Ho * Wo * C4) return;
  const int c = i % C4, p = i / C4, oy = p / Wo, ox = p - oy * Wo;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int y = 2 * oy + dy, x = 2 * ox + dx;
      if (y < 0 || y >= H || x < 0 || x >= W) continue;
      const float4 v = in[(((size_t)n * H + y) * W + x) * C4 + c];
      m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
    }
  out[(size_t)n * Ho * Wo * C4 + i] = m;
}

struct BbConv {                 // one BN-folded convolution on the tensor-core engine
  int cin = 0, cin2 = 0, cout = 0, k = 1;
  float* d_w = nullptr;         // fp16 hi/lo images, variants 0 / 1 / 2 (see fvp_debug_pack_tc16)
  const float* wtc16[3] = {nullptr, nullptr, nullptr};
  float* d_bias = nullptr;
};

}  // namespace

struct fvp_backbone {
  int device = 0, num_layers = 50, max_images = 0, max_h = 0, max_w = 0, num_sms = 148;
  bool bottleneck = true;
  std::string err;
  std::map<std::string, std::vector<float>> params;     // raw state_dict entries of the slice (others are accepted and dropped)
  bool ready = false;
  float *d_stem_w = nullptr, *d_stem_b = nullptr;
  std::vector<BbConv> convs;                             // layer1 in execution order
  float *d_stem = nullptr, *d_pool = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_a = nullptr, *d_b = nullptr;
};

static std::string g_bb_error;
static int bb_fail(fvp_backbone* bb, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (bb) bb->err = buf; else g_bb_error = buf;
  return code;
}

namespace {

int blocks_in_layer1(int num_layers) { return num_layers == 18 ? 2 : 3; }

bool needed(const fvp_backbone* bb, const std::string& name) {
  if (name.rfind("conv1.", 0) == 0 || name.rfind("bn1.", 0) == 0) return true;
  return name.rfind("layer1.", 0) == 0;
}

// y = s * conv + t of an eval-mode BatchNorm (eps 1e-5) after a bias-free convolution
bool bn_fold(const fvp_backbone* bb, const std::string& bn, int c, std::vector<double>& s, std::vector<double>& t) {
  auto g = bb->params.find(bn + ".weight"), b = bb->params.find(bn + ".bias"), m = bb->params.find(bn + ".running_mean"),
       v = bb->params.find(bn + ".running_var");
  if (g == bb->params.end() || b == bb->params.end() || m == bb->params.end() || v == bb->params.end()) return false;
  if ((int)g->second.size() != c) return false;
  s.resize(c); t.resize(c);
  for (int i = 0; i < c; ++i) {
    s[i] = (double)g->second[i] / std::sqrt((double)v->second[i] + 1e-5);
    t[i] = (double)b->second[i] - (double)m->second[i] * s[i];
  }
  return true;
}

// GEMM rows [(k*k*cinP + cin2P)][cout] of conv `key` (+ the 1x1 `skip` conv as extra rows), BN folded; bias = t (+ t_skip)
int pack_conv(fvp_backbone* bb, const std::string& key, const std::string& bn, int cin, int cout, int k, const std::string& skip,
              const std::string& skip_bn, int cin2, BbConv& out) {
  auto w = bb->params.find(key + ".weight");
  if (w == bb->params.end() || (int64_t)w->second.size() != (int64_t)cout * cin * k * k) return bb_fail(bb, FVP_E_STATE, "parameter '%s.weight' missing or misshapen", key.c_str());
  std::vector<double> s, t;
  if (!bn_fold(bb, bn, cout, s, t)) return bb_fail(bb, FVP_E_STATE, "BatchNorm '%s' missing or misshapen", bn.c_str());
  const int cinP = fvp_round_up(cin, 16), cin2P = cin2 ? fvp_round_up(cin2, 16) : 0, taps = k * k;
  std::vector<float> rows((size_t)(taps * cinP + cin2P) * cout, 0.f), bias(cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tp = 0; tp < taps; ++tp)
        rows[(size_t)(tp * cinP + ci) * cout + co] = (float)((double)w->second[((size_t)co * cin + ci) * taps + tp] * s[co]);
  std::vector<double> b(t);
  if (cin2) {
    auto w2 = bb->params.find(skip + ".weight");
    std::vector<double> s2, t2;
    if (w2 == bb->params.end() || (int64_t)w2->second.size() != (int64_t)cout * cin2 || !bn_fold(bb, skip_bn, cout, s2, t2))
      return bb_fail(bb, FVP_E_STATE, "downsample branch '%s' missing or misshapen", skip.c_str());
    for (int co = 0; co < cout; ++co) {
      for (int ci = 0; ci < cin2; ++ci) rows[(size_t)(taps * cinP + ci) * cout + co] = (float)((double)w2->second[(size_t)co * cin2 + ci] * s2[co]);
      b[co] += t2[co];
    }
  }
  for (int co = 0; co < cout; ++co) bias[co] = (float)b[co];
  for (float v : rows)
    if (!(std::fabs(v) < 65504.0f)) return bb_fail(bb, FVP_E_RANGE, "BN-folded weights of '%s' leave the fp16 range of the tensor-core engine", key.c_str());
  // weight images for N tiles of up to 128 / 32 / 64 columns (only where they differ), one allocation
  const int npad = fvp_round_up(cout, 16);
  std::vector<unsigned short> img;
  size_t off[3] = {(size_t)-1, (size_t)-1, (size_t)-1};
  for (int v = 0; v < 3; ++v) {
    if (!(v == 0 || (v == 1 && npad > 32) || (v == 2 && npad > 64))) continue;
    long long nh = 0;
    fvp_debug_pack_tc16(rows.data(), cin, cin2, cout, k, v, 32, nullptr, 0, &nh);
    const size_t at = (img.size() + 127) & ~(size_t)127;          // 256-byte aligned
    img.resize(at + (size_t)nh);
    if (fvp_debug_pack_tc16(rows.data(), cin, cin2, cout, k, v, 32, img.data() + at, nh, &nh) != 0) return bb_fail(bb, FVP_E_INVALID, "weight packing failed for '%s'", key.c_str());
    off[v] = at;
  }
  out.cin = cin; out.cin2 = cin2; out.cout = cout; out.k = k;
  if (cudaMalloc((void**)&out.d_w, img.size() * 2) != cudaSuccess || cudaMalloc((void**)&out.d_bias, (size_t)fvp_round_up(cout, 4) * 4) != cudaSuccess)
    return bb_fail(bb, FVP_E_CUDA, "allocation failed");
  cudaMemcpy(out.d_w, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(out.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  for (int v = 0; v < 3; ++v) out.wtc16[v] = off[v] == (size_t)-1 ? nullptr : (const float*)((const unsigned short*)out.d_w + off[v]);
  return FVP_OK;
}

void run_conv(const fvp_backbone* bb, const BbConv& c, const float* in, const float* in2, const float* res, float* out, int n, int H, int W,
              cudaStream_t st) {
  FvpConvArgs a;
  a.in = in; a.H = H; a.W = W; a.Cin = c.cin; a.in2 = in2; a.Cin2 = c.cin2; a.w = nullptr; a.bias = c.d_bias; a.out = out;
  a.CoutP = c.cout; a.CoutS = c.cout; a.CoutReal = c.cout; a.res = res; a.res_mode = res ? 1 : 0; a.relu = 1; a.ksize = c.k;
  a.upsample = 0; a.nchw = 0; a.n = n; a.valid = nullptr; a.fmt = 0;
  const FvpLaunchEnv env{bb->num_sms, 2, nullptr, nullptr, 0};
  fvp_launch_conv_tc(a, c.wtc16, 1, env, st);
}

}  // namespace

extern "C" {

const char* fvp_backbone_last_error(const fvp_backbone* bb) { return bb ? bb->err.c_str() : g_bb_error.c_str(); }

int fvp_backbone_create(int num_layers, int max_images, int max_h, int max_w, int device, fvp_backbone** out) {
  if (!out) return bb_fail(nullptr, FVP_E_INVALID, "null argument");
  *out = nullptr;
  if (num_layers != 18 && num_layers != 34 && num_layers != 50 && num_layers != 101 && num_layers != 152)
    return bb_fail(nullptr, FVP_E_INVALID, "RESNET.NUM_LAYERS must be 18/34/50/101/152 (resnet.py:204-208)");
  if (max_images < 1 || max_h < 32 || max_w < 32 || max_h % 4 || max_w % 4) return bb_fail(nullptr, FVP_E_INVALID, "image size must be a multiple of 4, >= 32");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bb_fail(nullptr, FVP_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return bb_fail(nullptr, FVP_E_CUDA, "libfvp_b200 is built for sm_100a only");
  e = fvp_conv_tc_init_device();
  if (e != cudaSuccess) return bb_fail(nullptr, FVP_E_CUDA, "kernel attribute setup: %s", cudaGetErrorString(e));
  fvp_backbone* bb = new fvp_backbone();
  bb->device = device; bb->num_layers = num_layers; bb->max_images = max_images; bb->max_h = max_h; bb->max_w = max_w;
  bb->bottleneck = num_layers >= 50;
  bb->num_sms = prop.multiProcessorCount;
  const size_t half = (size_t)max_images * (max_h / 2) * (max_w / 2), quarter = (size_t)max_images * (max_h / 4) * (max_w / 4);
  const int cexp = bb->bottleneck ? 256 : 64;
  bool ok = cudaMalloc((void**)&bb->d_stem, half * 64 * 4) == cudaSuccess && cudaMalloc((void**)&bb->d_pool, quarter * 64 * 4) == cudaSuccess &&
            cudaMalloc((void**)&bb->d_t1, quarter * 64 * 4) == cudaSuccess && cudaMalloc((void**)&bb->d_t2, quarter * 64 * 4) == cudaSuccess &&
            cudaMalloc((void**)&bb->d_a, quarter * cexp * 4) == cudaSuccess && cudaMalloc((void**)&bb->d_b, quarter * cexp * 4) == cudaSuccess;
  if (!ok) {
    fvp_backbone_destroy(bb);
    return bb_fail(nullptr, FVP_E_CUDA, "workspace allocation failed");
  }
  *out = bb;
  return FVP_OK;
}

void fvp_backbone_destroy(fvp_backbone* bb) {
  if (!bb) return;
  cudaSetDevice(bb->device);
  cudaDeviceSynchronize();
  void* p[] = {bb->d_stem_w, bb->d_stem_b, bb->d_stem, bb->d_pool, bb->d_t1, bb->d_t2, bb->d_a, bb->d_b};
  for (void* q : p)
    if (q) cudaFree(q);
  for (BbConv& c : bb->convs) {
    if (c.d_w) cudaFree(c.d_w);
    if (c.d_bias) cudaFree(c.d_bias);
  }
  delete bb;
}

int fvp_backbone_set_param(fvp_backbone* bb, const char* name, const float* h_data, int64_t numel) {
  if (!bb || !name) return FVP_E_INVALID;
  const std::string k(name);
  if (!needed(bb, k) || k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return FVP_OK;   // later layers: accepted, not built yet
  if (!h_data || numel < 1) return bb_fail(bb, FVP_E_INVALID, "null data for %s", name);
  bb->params[k].assign(h_data, h_data + numel);
  bb->ready = false;
  return FVP_OK;
}

int fvp_backbone_finalize(fvp_backbone* bb) {
  if (!bb) return FVP_E_INVALID;
  cudaSetDevice(bb->device);
  cudaDeviceSynchronize();
  for (BbConv& c : bb->convs) {
    if (c.d_w) cudaFree(c.d_w);
    if (c.d_bias) cudaFree(c.d_bias);
  }
  bb->convs.clear();
  // stem: [co][3][7][7] -> rows [(c, dy, dx)][64], BN folded
  auto w = bb->params.find("conv1.weight");
  std::vector<double> s, t;
  if (w == bb->params.end() || w->second.size() != 64u * 3 * 49 || !bn_fold(bb, "bn1", 64, s, t)) return bb_fail(bb, FVP_E_STATE, "stem parameters (conv1 / bn1) missing or misshapen");
  std::vector<float> sw(147 * 64), sb(64);
  for (int co = 0; co < 64; ++co) {
    for (int r = 0; r < 147; ++r) sw[(size_t)r * 64 + co] = (float)((double)w->second[(size_t)co * 147 + r] * s[co]);
    sb[co] = (float)t[co];
  }
  if (!bb->d_stem_w && (cudaMalloc((void**)&bb->d_stem_w, sw.size() * 4) != cudaSuccess || cudaMalloc((void**)&bb->d_stem_b, 256) != cudaSuccess))
    return bb_fail(bb, FVP_E_CUDA, "allocation failed");
  cudaMemcpy(bb->d_stem_w, sw.data(), sw.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(bb->d_stem_b, sb.data(), 256, cudaMemcpyHostToDevice);
  // layer1
  const int nb = blocks_in_layer1(bb->num_layers);
  for (int b = 0; b < nb; ++b) {
    const std::string p = "layer1." + std::to_string(b);
    int rc;
    if (bb->bottleneck) {
      const int cin = b == 0 ? 64 : 256;
      bb->convs.emplace_back(); rc = pack_conv(bb, p + ".conv1", p + ".bn1", cin, 64, 1, "", "", 0, bb->convs.back()); if (rc) return rc;
      bb->convs.emplace_back(); rc = pack_conv(bb, p + ".conv2", p + ".bn2", 64, 64, 3, "", "", 0, bb->convs.back()); if (rc) return rc;
      bb->convs.emplace_back();
      rc = b == 0 ? pack_conv(bb, p + ".conv3", p + ".bn3", 64, 256, 1, p + ".downsample.0", p + ".downsample.1", 64, bb->convs.back())
                  : pack_conv(bb, p + ".conv3", p + ".bn3", 64, 256, 1, "", "", 0, bb->convs.back());
      if (rc) return rc;
    } else {
      bb->convs.emplace_back(); rc = pack_conv(bb, p + ".conv1", p + ".bn1", 64, 64, 3, "", "", 0, bb->convs.back()); if (rc) return rc;
      bb->convs.emplace_back(); rc = pack_conv(bb, p + ".conv2", p + ".bn2", 64, 64, 3, "", "", 0, bb->convs.back()); if (rc) return rc;
    }
  }
  bb->ready = true;
  return FVP_OK;
}

int fvp_backbone_forward_slice(fvp_backbone* bb, const float* d_images, int n, int h, int w, int stage, float* d_out, uintptr_t stream) {
  if (!bb || !d_images || !d_out) return bb_fail(bb, FVP_E_INVALID, "null argument");
  if (!bb->ready) return bb_fail(bb, FVP_E_STATE, "fvp_backbone_finalize has not been called");
  const int nb = blocks_in_layer1(bb->num_layers);
  if (n < 1 || n > bb->max_images || h < 32 || w < 32 || h > bb->max_h || w > bb->max_w || h % 4 || w % 4)
    return bb_fail(bb, FVP_E_INVALID, "images %d x %dx%d outside what the backbone was created for (%d x %dx%d, multiples of 4)", n, h, w, bb->max_images, bb->max_h, bb->max_w);
  if (stage < 0 || stage > nb) return bb_fail(bb, FVP_E_INVALID, "stage %d outside [0, %d] (later layers are not built yet)", stage, nb);
  cudaSetDevice(bb->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int ho = h / 2, wo = w / 2, H = h / 4, W = w / 4;
  k_stem7x7s2<<<dim3(fvp_cdiv(wo, ST_T), fvp_cdiv(ho, ST_T), n), 256, 0, st>>>(d_images, bb->d_stem_w, bb->d_stem_b, bb->d_stem, h, w, ho, wo);
  k_maxpool3s2<<<dim3(fvp_cdiv(H * W * 16, 256), n), 256, 0, st>>>((const float4*)bb->d_stem, (float4*)bb->d_pool, ho, wo, H, W, 16);
  const float* x = bb->d_pool;
  int cx = 64;
  float* pp[2] = {bb->d_a, bb->d_b};
  for (int b = 0; b < stage; ++b) {
    float* y = pp[b & 1];
    if (bb->bottleneck) {
      const BbConv *c1 = &bb->convs[3 * b], *c2 = c1 + 1, *c3 = c1 + 2;
      run_conv(bb, *c1, x, nullptr, nullptr, bb->d_t1, n, H, W, st);
      run_conv(bb, *c2, bb->d_t1, nullptr, nullptr, bb->d_t2, n, H, W, st);
      // block 0: conv3(y) + downsample(x) in one GEMM (second K segment), ReLU; later blocks: identity residual in the epilogue
      run_conv(bb, *c3, bb->d_t2, b == 0 ? x : nullptr, b == 0 ? nullptr : x, y, n, H, W, st);
      cx = 256;
    } else {
      const BbConv *c1 = &bb->convs[2 * b], *c2 = c1 + 1;
      run_conv(bb, *c1, x, nullptr, nullptr, bb->d_t1, n, H, W, st);
      run_conv(bb, *c2, bb->d_t1, nullptr, x, y, n, H, W, st);
    }
    x = y;
  }
  fvp_launch_nhwc_to_nchw(x, d_out, n, H * W, cx, cx, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bb_fail(bb, FVP_E_CUDA, "CUDA error: %s", cudaGetErrorString(e));
  return FVP_OK;
}

}  // extern "C"
