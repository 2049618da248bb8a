// fvp_backbone.cu - N2: the PoseResNet backbone (lib/models/resnet.py:98-201), the step in front of the hot
// path when TEST_HEATMAP_SRC = 'image' (faster_voxelpose.py:36-38: heat maps = backbone(views[:, c])).
//
//   stem      conv 7x7 stride 2 (3 -> 64, no bias) + BatchNorm(eval) + ReLU      resnet.py:103-107,189-191
//             k_stem7x7s2: exact-fp32 CUDA-core kernel (3 input channels = 2 % of the backbone's MACs), NCHW images in,
//             NHWC out, BN folded into the weights
//   max-pool  3x3 stride 2 pad 1                                                  resnet.py:108,192
//   layer1-4  Bottleneck (1x1, 3x3 [stride on the 3x3], 1x1 x4, + 1x1 downsample of the block input in a stage's first
//             block; out += residual; ReLU, resnet.py:57-95,118-146) or BasicBlock (resnet.py:22-54) on the tcgen05 engine
//             of fvp_conv_tc.cu (fp16 hi/lo split, fp32 accumulation) in its TMA-fed form: from the max-pool on, every
//             activation is a split tensor (fp16 hi plane + scaled-lo plane) written once by the producing epilogue and
//             fetched by cp.async.bulk.tensor; BN folded, the downsample conv fused as the second K
//             segment of the block's last conv, the identity residual added in its epilogue; output channels beyond 256
//             (the engine's bias table) run as 256-column launches into one NHWC tensor.
//             Stride 2 (first version): a stride-2 3x3 is the stride-1 convolution kept at the even positions
//             (k_decimate2 after it: 4x the MACs on 3 of 53 layers, ~9 % of the network), a stride-2 1x1 reads the
//             decimated input (exact).
//   deconv    ConvTranspose2d(4, stride 2, pad 1) + BN + ReLU (resnet.py:148-186): every output parity (a, b) is a 2x2
//             convolution of the input, written here as ONE 3x3 convolution to 4 x Cout columns (rows {-1,0} or {0,+1}
//             of the taps non-zero per parity) with the engine's pixel-shuffle epilogue (2.25x the MACs, first version)
//   final     1x1 convolution with bias to the J heat maps, planar fp32 output        resnet.py:175-186,198-199
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_fp16.h>

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

// MaxPool2d(3, 2, 1) on NHWC, 4 channels per thread (padding = -inf, i.e. ignored).  The result is written as a SPLIT
// tensor (fp16 hi plane, then the scaled-lo plane: the operand form of the tensor-core engine, see fvp_conv_tc.cu) -
// everything downstream of the max-pool is fed by TMA.
__global__ void __launch_bounds__(256) k_maxpool3s2(const float4* __restrict__ in, __half* __restrict__ out, int H, int W, int Ho, int Wo,
                                                    int C4, size_t plane) {
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
  const __half2 h01 = __floats2half2_rn(m.x, m.y), h23 = __floats2half2_rn(m.z, m.w);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((m.x - f01.x) * 2048.0f, (m.y - f01.y) * 2048.0f);
  const __half2 l23 = __floats2half2_rn((m.z - f23.x) * 2048.0f, (m.w - f23.y) * 2048.0f);
  __half* o = out + ((size_t)n * Ho * Wo * C4 + i) * 4;
  *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(o + plane) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// keep the even positions of a split NHWC tensor: [n][H][W][C] -> [n][ceil(H/2)][ceil(W/2)][C], 8 channels (16 bytes of
// either plane) per thread
__global__ void __launch_bounds__(256) k_decimate2(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int Ho, int Wo, int C8,
                                                   size_t plane_in, size_t plane_out) {
  const int n = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Ho * Wo * C8) return;
  const int c = i % C8, p = i / C8, oy = p / Wo, ox = p - oy * Wo;
  const size_t src = (((size_t)n * H + 2 * oy) * W + 2 * ox) * C8 + c, dst = (size_t)n * Ho * Wo * C8 + i;
  out[dst] = in[src];
  out[plane_out + dst] = in[plane_in + src];
}

struct BbChunk {                // <= 256 GEMM columns of one convolution: its own weight images and bias slice
  int cols = 0;                 // GEMM columns of this launch (CoutP)
  int ch0 = 0;                  // first output channel it writes
  float* d_w = nullptr;
  const float* wtc16[3] = {nullptr, nullptr, nullptr};
  float* d_bias = nullptr;
};
struct BbConv {                 // one BN-folded convolution on the tensor-core engine
  int cin = 0, cin2 = 0, cout = 0, k = 1;
  int upsample = 0;             // transposed 4x4 stride-2 convolution written as 3x3 + pixel shuffle
  std::vector<BbChunk> chunks;
};
struct BbBlock {                // one residual block of the program
  int c1 = -1, c2 = -1, c3 = -1;   // indices into convs (c3 = -1 for BasicBlock)
  int stride = 1;
  bool downsample = false;
  int cin = 0, planes = 0, cout = 0;
};

}  // namespace

struct fvp_backbone {
  int device = 0, num_layers = 50, num_joints = 15, max_images = 0, max_h = 0, max_w = 0, num_sms = 148;
  bool bottleneck = true;
  std::string err;
  std::map<std::string, std::vector<float>> params;     // raw state_dict entries
  bool ready = false;
  int launch_error = 0;
  float *d_stem_w = nullptr, *d_stem_b = nullptr;
  std::vector<BbConv> convs;
  std::vector<BbBlock> blocks;
  int deconv[3] = {-1, -1, -1};
  int final_conv = -1;
  float* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // NHWC scratch units (largest tensor of the net each)
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

const int* layer_blocks(int num_layers) {
  static const int r18[4] = {2, 2, 2, 2}, r34[4] = {3, 4, 6, 3}, r101[4] = {3, 4, 23, 3}, r152[4] = {3, 8, 36, 3};
  return num_layers == 18 ? r18 : (num_layers == 101 ? r101 : (num_layers == 152 ? r152 : r34));   // 34 and 50 share [3,4,6,3]
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

// upload one <= 256-column slice of a GEMM matrix rows[(k*k*cinP + cin2P)][ncols_total] (columns col0 .. col0+cols)
int upload_chunk(fvp_backbone* bb, const std::vector<float>& rows, int nrows, int ncols_total, int col0, int cols, const std::vector<float>& bias,
                 int cin, int cin2, int k, int ch0, BbChunk& out, const char* what) {
  std::vector<float> sub((size_t)nrows * cols);
  for (int r = 0; r < nrows; ++r) memcpy(&sub[(size_t)r * cols], &rows[(size_t)r * ncols_total + col0], (size_t)cols * 4);
  for (float v : sub)
    if (!(std::fabs(v) < 65504.0f)) return bb_fail(bb, FVP_E_RANGE, "BN-folded weights of '%s' leave the fp16 range of the tensor-core engine", what);
  const int npad = fvp_round_up(cols, 16);
  std::vector<unsigned short> img;
  size_t off[3] = {(size_t)-1, (size_t)-1, (size_t)-1};
  for (int v = 0; v < 3; ++v) {
    if (!(v == 0 || (v == 1 && npad > 32) || (v == 2 && npad > 64))) continue;
    long long nh = 0;
    fvp_debug_pack_tc16(sub.data(), cin, cin2, cols, k, v, 32, nullptr, 0, &nh);
    const size_t at = (img.size() + 127) & ~(size_t)127;          // 256-byte aligned
    img.resize(at + (size_t)nh);
    if (fvp_debug_pack_tc16(sub.data(), cin, cin2, cols, k, v, 32, img.data() + at, nh, &nh) != 0) return bb_fail(bb, FVP_E_INVALID, "weight packing failed for '%s'", what);
    off[v] = at;
  }
  out.cols = cols; out.ch0 = ch0;
  std::vector<float> b4((size_t)fvp_round_up(cols, 4), 0.f);
  for (int i = 0; i < cols; ++i) b4[i] = bias[col0 + i];
  if (cudaMalloc((void**)&out.d_w, img.size() * 2) != cudaSuccess || cudaMalloc((void**)&out.d_bias, b4.size() * 4) != cudaSuccess)
    return bb_fail(bb, FVP_E_CUDA, "allocation failed");
  cudaMemcpy(out.d_w, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(out.d_bias, b4.data(), b4.size() * 4, cudaMemcpyHostToDevice);
  for (int v = 0; v < 3; ++v) out.wtc16[v] = off[v] == (size_t)-1 ? nullptr : (const float*)((const unsigned short*)out.d_w + off[v]);
  return FVP_OK;
}

// regular convolution `key` (+ the 1x1 `skip` conv as extra GEMM rows), BN folded (bn == "": plain bias `key.bias`)
int pack_conv(fvp_backbone* bb, const std::string& key, const std::string& bn, int cin, int cout, int k, const std::string& skip,
              const std::string& skip_bn, int cin2, BbConv& out) {
  auto w = bb->params.find(key + ".weight");
  if (w == bb->params.end() || (int64_t)w->second.size() != (int64_t)cout * cin * k * k) return bb_fail(bb, FVP_E_STATE, "parameter '%s.weight' missing or misshapen", key.c_str());
  std::vector<double> s(cout, 1.0), t(cout, 0.0);
  if (!bn.empty()) {
    if (!bn_fold(bb, bn, cout, s, t)) return bb_fail(bb, FVP_E_STATE, "BatchNorm '%s' missing or misshapen", bn.c_str());
  } else {
    auto b = bb->params.find(key + ".bias");
    if (b == bb->params.end() || (int)b->second.size() != cout) return bb_fail(bb, FVP_E_STATE, "parameter '%s.bias' missing or misshapen", key.c_str());
    for (int i = 0; i < cout; ++i) t[i] = b->second[i];
  }
  const int cinP = fvp_round_up(cin, 16), cin2P = cin2 ? fvp_round_up(cin2, 16) : 0, taps = k * k;
  const int coutP = fvp_round_up(cout, 4), nrows = taps * cinP + cin2P;
  std::vector<float> rows((size_t)nrows * coutP, 0.f), bias(coutP, 0.f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tp = 0; tp < taps; ++tp)
        rows[(size_t)(tp * cinP + ci) * coutP + co] = (float)((double)w->second[((size_t)co * cin + ci) * taps + tp] * s[co]);
  std::vector<double> b(t);
  if (cin2) {
    auto w2 = bb->params.find(skip + ".weight");
    std::vector<double> s2, t2;
    if (w2 == bb->params.end() || (int64_t)w2->second.size() != (int64_t)cout * cin2 || !bn_fold(bb, skip_bn, cout, s2, t2))
      return bb_fail(bb, FVP_E_STATE, "downsample branch '%s' missing or misshapen", skip.c_str());
    for (int co = 0; co < cout; ++co) {
      for (int ci = 0; ci < cin2; ++ci) rows[(size_t)(taps * cinP + ci) * coutP + co] = (float)((double)w2->second[(size_t)co * cin2 + ci] * s2[co]);
      b[co] += t2[co];
    }
  }
  for (int co = 0; co < cout; ++co) bias[co] = (float)b[co];
  out.cin = cin; out.cin2 = cin2; out.cout = cout; out.k = k; out.upsample = 0;
  for (int c0 = 0; c0 < coutP; c0 += 256) {
    out.chunks.emplace_back();
    const int rc = upload_chunk(bb, rows, nrows, coutP, c0, coutP - c0 < 256 ? coutP - c0 : 256, bias, cin, cin2, k, c0, out.chunks.back(), key.c_str());
    if (rc) return rc;
  }
  return FVP_OK;
}

// ConvTranspose2d(4, 2, 1) [Cin][Cout][4][4] + BN as a 3x3 convolution to 4 x Cout columns (q = a*2 + b = output parity):
// out[2y+a, 2x+b] = sum over input (y + r, x + c), r / c in {-1, 0, +1}, of W[ky = a + 1 - 2r][kx = b + 1 - 2c] where 0 <= ky, kx <= 3.
// Launches cover 64 output channels x 4 parities each (column q*64 + c of the launch = the engine's pixel-shuffle order).
int pack_deconv(fvp_backbone* bb, const std::string& key, const std::string& bn, int cin, int cout, BbConv& out) {
  auto w = bb->params.find(key + ".weight");
  if (w == bb->params.end() || (int64_t)w->second.size() != (int64_t)cin * cout * 16) return bb_fail(bb, FVP_E_STATE, "parameter '%s.weight' missing or misshapen", key.c_str());
  if (cout % 64) return bb_fail(bb, FVP_E_INVALID, "NUM_DECONV_FILTERS must be multiples of 64");
  std::vector<double> s, t;
  if (!bn_fold(bb, bn, cout, s, t)) return bb_fail(bb, FVP_E_STATE, "BatchNorm '%s' missing or misshapen", bn.c_str());
  if (bb->params.count(key + ".bias")) {                       // DECONV_WITH_BIAS: a bias before the BN folds into its shift
    const std::vector<float>& b = bb->params[key + ".bias"];
    for (int i = 0; i < cout; ++i) t[i] += (double)b[i] * s[i];
  }
  const int cinP = fvp_round_up(cin, 16), nrows = 9 * cinP;
  out.cin = cin; out.cin2 = 0; out.cout = cout; out.k = 3; out.upsample = 1;
  for (int c0 = 0; c0 < cout; c0 += 64) {
    std::vector<float> rows((size_t)nrows * 256, 0.f), bias(256, 0.f);
    for (int q = 0; q < 4; ++q) {
      const int a = q >> 1, b = q & 1;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const int ky = a + 1 - 2 * (dy - 1), kx = b + 1 - 2 * (dx - 1);
          if (ky < 0 || ky > 3 || kx < 0 || kx > 3) continue;
          for (int c = 0; c < 64; ++c) {
            const int co = c0 + c;
            for (int ci = 0; ci < cin; ++ci)
              rows[(size_t)((dy * 3 + dx) * cinP + ci) * 256 + q * 64 + c] = (float)((double)w->second[(((size_t)ci * cout + co) * 4 + ky) * 4 + kx] * s[co]);
          }
        }
      for (int c = 0; c < 64; ++c) bias[q * 64 + c] = (float)t[c0 + c];
    }
    out.chunks.emplace_back();
    const int rc = upload_chunk(bb, rows, nrows, 256, 0, 256, bias, cin, 0, 3, c0, out.chunks.back(), key.c_str());
    if (rc) return rc;
  }
  return FVP_OK;
}

// One convolution = one launch per <= 256-column chunk into the same NHWC tensor (channel stride = all output channels).
// All activations are SPLIT tensors (fp16 hi plane + scaled-lo plane) fetched by TMA; only the heat maps leave as fp32 NCHW.
void run_conv(const fvp_backbone* bb, const BbConv& c, const float* in, const float* in2, const float* res, float* out, int n, int H, int W,
              int relu, cudaStream_t st, float* nchw_out = nullptr, int cout_real = 0) {
  for (const BbChunk& ch : c.chunks) {
    FvpConvArgs a;
    a.in = in; a.H = H; a.W = W; a.Cin = c.cin; a.in2 = in2; a.Cin2 = c.cin2; a.w = nullptr; a.bias = ch.d_bias;
    // channel offsets of a chunk are in ELEMENTS of the split planes (halves)
    a.out = nchw_out ? nchw_out : reinterpret_cast<float*>(reinterpret_cast<__half*>(out) + ch.ch0);
    a.CoutP = ch.cols; a.CoutS = c.cout; a.CoutReal = cout_real;
    a.res = res ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(res) + ch.ch0) : nullptr;
    a.res_mode = res ? 1 : 0; a.relu = relu; a.ksize = c.k;
    a.upsample = c.upsample; a.nchw = nchw_out ? 1 : 0; a.n = n; a.valid = nullptr;
    a.fmt = FVP_FMT_IN_SPLIT | (nchw_out ? 0 : FVP_FMT_OUT_SPLIT) | (res ? FVP_FMT_RES_SPLIT : 0);
    const FvpLaunchEnv env{bb->num_sms, 2, nullptr, nullptr, 0, const_cast<int*>(&bb->launch_error)};
    fvp_launch_conv_tc(a, ch.wtc16, 1, env, st);
  }
}
void decimate(const float* in, float* out, int n, int H, int W, int C, cudaStream_t st) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  k_decimate2<<<dim3(fvp_cdiv(Ho * Wo * (C / 8), 256), n), 256, 0, st>>>((const uint4*)in, (uint4*)out, H, W, Ho, Wo, C / 8,
                                                                        (size_t)n * H * W * (C / 8), (size_t)n * Ho * Wo * (C / 8));
}

void free_convs(fvp_backbone* bb) {
  for (BbConv& c : bb->convs)
    for (BbChunk& ch : c.chunks) {
      if (ch.d_w) cudaFree(ch.d_w);
      if (ch.d_bias) cudaFree(ch.d_bias);
    }
  bb->convs.clear();
  bb->blocks.clear();
}

}  // namespace

extern "C" {

const char* fvp_backbone_last_error(const fvp_backbone* bb) { return bb ? bb->err.c_str() : g_bb_error.c_str(); }

int fvp_backbone_create(int num_layers, int num_joints, int max_images, int max_h, int max_w, int device, fvp_backbone** out) {
  if (!out) return bb_fail(nullptr, FVP_E_INVALID, "null argument");
  *out = nullptr;
  if (num_layers != 18 && num_layers != 34 && num_layers != 50 && num_layers != 101 && num_layers != 152)
    return bb_fail(nullptr, FVP_E_INVALID, "RESNET.NUM_LAYERS must be 18/34/50/101/152 (resnet.py:204-208)");
  if (num_joints < 1 || num_joints > 64) return bb_fail(nullptr, FVP_E_INVALID, "num_joints must be 1..64");
  if (max_images < 1 || max_h < 32 || max_w < 32 || max_h % 32 || max_w % 32) return bb_fail(nullptr, FVP_E_INVALID, "image size must be a multiple of 32, >= 32");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bb_fail(nullptr, FVP_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return bb_fail(nullptr, FVP_E_CUDA, "libfvp_b200 is built for sm_100a only");
  e = fvp_conv_tc_init_device();
  if (e != cudaSuccess) return bb_fail(nullptr, FVP_E_CUDA, "kernel attribute setup: %s", cudaGetErrorString(e));
  fvp_backbone* bb = new fvp_backbone();
  bb->device = device; bb->num_layers = num_layers; bb->num_joints = num_joints;
  bb->max_images = max_images; bb->max_h = max_h; bb->max_w = max_w;
  bb->bottleneck = num_layers >= 50;
  bb->num_sms = prop.multiProcessorCount;
  // largest NHWC tensor of the network: the stem output (h/2 x w/2 x 64) = layer1's output (h/4 x w/4 x 256) = a deconv output
  const size_t unit = (size_t)max_images * (max_h / 2) * (max_w / 2) * 64;
  for (int i = 0; i < 6; ++i)
    if (cudaMalloc((void**)&bb->buf[i], unit * 4) != cudaSuccess) {
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
  if (bb->d_stem_w) cudaFree(bb->d_stem_w);
  if (bb->d_stem_b) cudaFree(bb->d_stem_b);
  for (float* q : bb->buf)
    if (q) cudaFree(q);
  free_convs(bb);
  delete bb;
}

int fvp_backbone_set_param(fvp_backbone* bb, const char* name, const float* h_data, int64_t numel) {
  if (!bb || !name) return FVP_E_INVALID;
  const std::string k(name);
  if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return FVP_OK;
  if (!h_data || numel < 1) return bb_fail(bb, FVP_E_INVALID, "null data for %s", name);
  bb->params[k].assign(h_data, h_data + numel);
  bb->ready = false;
  return FVP_OK;
}

int fvp_backbone_num_stages(const fvp_backbone* bb) { return bb ? (int)bb->blocks.size() + 5 : 0; }   // pool, blocks, 3 deconvs, heat maps

int fvp_backbone_finalize(fvp_backbone* bb) {
  if (!bb) return FVP_E_INVALID;
  cudaSetDevice(bb->device);
  cudaDeviceSynchronize();
  free_convs(bb);
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
  // residual stages (resnet.py:118-146)
  const int* nblk = layer_blocks(bb->num_layers);
  const int exp = bb->bottleneck ? 4 : 1;
  int inplanes = 64;
  auto add = [&](int& idx) -> BbConv& { bb->convs.emplace_back(); idx = (int)bb->convs.size() - 1; return bb->convs.back(); };
  for (int li = 1; li <= 4; ++li) {
    const int planes = 64 << (li - 1);
    for (int b = 0; b < nblk[li - 1]; ++b) {
      BbBlock blk;
      blk.stride = (b == 0 && li > 1) ? 2 : 1;
      blk.cin = inplanes; blk.planes = planes; blk.cout = planes * exp;
      blk.downsample = b == 0 && (blk.stride != 1 || inplanes != planes * exp);
      const std::string p = "layer" + std::to_string(li) + "." + std::to_string(b);
      int rc;
      if (bb->bottleneck) {
        rc = pack_conv(bb, p + ".conv1", p + ".bn1", inplanes, planes, 1, "", "", 0, add(blk.c1)); if (rc) return rc;
        rc = pack_conv(bb, p + ".conv2", p + ".bn2", planes, planes, 3, "", "", 0, add(blk.c2)); if (rc) return rc;
        rc = blk.downsample ? pack_conv(bb, p + ".conv3", p + ".bn3", planes, planes * 4, 1, p + ".downsample.0", p + ".downsample.1", inplanes, add(blk.c3))
                            : pack_conv(bb, p + ".conv3", p + ".bn3", planes, planes * 4, 1, "", "", 0, add(blk.c3));
        if (rc) return rc;
      } else {
        rc = pack_conv(bb, p + ".conv1", p + ".bn1", inplanes, planes, 3, "", "", 0, add(blk.c1)); if (rc) return rc;
        rc = blk.downsample ? pack_conv(bb, p + ".conv2", p + ".bn2", planes, planes, 3, p + ".downsample.0", p + ".downsample.1", inplanes, add(blk.c2))
                            : pack_conv(bb, p + ".conv2", p + ".bn2", planes, planes, 3, "", "", 0, add(blk.c2));
        if (rc) return rc;
      }
      bb->blocks.push_back(blk);
      inplanes = planes * exp;
    }
  }
  // three transposed convolutions to 256 channels (NUM_DECONV_FILTERS, kernel 4) and the final 1x1 (FINAL_CONV_KERNEL 1)
  for (int d = 0; d < 3; ++d) {
    const std::string key = "deconv_layers." + std::to_string(3 * d);
    auto dw = bb->params.find(key + ".weight");
    if (dw == bb->params.end() || dw->second.size() % ((size_t)inplanes * 16)) return bb_fail(bb, FVP_E_STATE, "parameter '%s.weight' missing or misshapen (4x4 transposed convolutions only)", key.c_str());
    const int cout = (int)(dw->second.size() / ((size_t)inplanes * 16));
    const int rc = pack_deconv(bb, key, "deconv_layers." + std::to_string(3 * d + 1), inplanes, cout, add(bb->deconv[d]));
    if (rc) return rc;
    inplanes = cout;
  }
  {
    auto fw = bb->params.find("final_layer.weight");
    if (fw == bb->params.end() || (int64_t)fw->second.size() != (int64_t)bb->num_joints * inplanes) return bb_fail(bb, FVP_E_STATE, "final_layer.weight missing or misshapen (1x1 final convolution only)");
    const int rc = pack_conv(bb, "final_layer", "", inplanes, bb->num_joints, 1, "", "", 0, add(bb->final_conv));
    if (rc) return rc;
  }
  bb->ready = true;
  return FVP_OK;
}

// stage: 0 = after the max-pool; 1..B = after residual block stage-1 (B = number of blocks); B+1..B+3 = after a transposed
// convolution; B+4 (= fvp_backbone_num_stages - 1) = the heat maps.  d_out: NCHW fp32 of that tap.
int fvp_backbone_forward_slice(fvp_backbone* bb, const float* d_images, int n, int h, int w, int stage, float* d_out, uintptr_t stream) {
  if (!bb || !d_images || !d_out) return bb_fail(bb, FVP_E_INVALID, "null argument");
  if (!bb->ready) return bb_fail(bb, FVP_E_STATE, "fvp_backbone_finalize has not been called");
  const int nb = (int)bb->blocks.size(), last = nb + 4;
  if (n < 1 || n > bb->max_images || h < 32 || w < 32 || h > bb->max_h || w > bb->max_w || h % 32 || w % 32)
    return bb_fail(bb, FVP_E_INVALID, "images %d x %dx%d outside what the backbone was created for (%d x %dx%d, multiples of 32)", n, h, w, bb->max_images, bb->max_h, bb->max_w);
  if (stage < 0 || stage > last) return bb_fail(bb, FVP_E_INVALID, "stage %d outside [0, %d]", stage, last);
  cudaSetDevice(bb->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int ho = h / 2, wo = w / 2;
  int H = h / 4, W = w / 4;
  float *X = bb->buf[0], *Y = bb->buf[1], *T1 = bb->buf[2], *T2 = bb->buf[3], *T3 = bb->buf[4], *XS = bb->buf[5];
  k_stem7x7s2<<<dim3(fvp_cdiv(wo, ST_T), fvp_cdiv(ho, ST_T), n), 256, 0, st>>>(d_images, bb->d_stem_w, bb->d_stem_b, Y, h, w, ho, wo);
  k_maxpool3s2<<<dim3(fvp_cdiv(H * W * 16, 256), n), 256, 0, st>>>((const float4*)Y, (__half*)X, ho, wo, H, W, 16, (size_t)n * H * W * 64);
  int cx = 64;
  for (int b = 0; b < nb && b < stage; ++b) {
    const BbBlock& k = bb->blocks[b];
    const int Ho = k.stride == 2 ? (H + 1) / 2 : H, Wo = k.stride == 2 ? (W + 1) / 2 : W;
    const float* xs = X;                                   // the block input at the block's OUTPUT resolution (1x1 downsample / identity)
    if (k.stride == 2 && k.downsample) { decimate(X, XS, n, H, W, k.cin, st); xs = XS; }
    if (bb->bottleneck) {
      run_conv(bb, bb->convs[k.c1], X, nullptr, nullptr, T1, n, H, W, 1, st);
      const float* t2 = T2;
      run_conv(bb, bb->convs[k.c2], T1, nullptr, nullptr, T2, n, H, W, 1, st);          // stride 2: stride-1 result ...
      if (k.stride == 2) { decimate(T2, T3, n, H, W, k.planes, st); t2 = T3; }         // ... kept at the even positions
      run_conv(bb, bb->convs[k.c3], t2, k.downsample ? xs : nullptr, k.downsample ? nullptr : X, Y, n, Ho, Wo, 1, st);
    } else {
      const float* t1 = T1;
      run_conv(bb, bb->convs[k.c1], X, nullptr, nullptr, T1, n, H, W, 1, st);
      if (k.stride == 2) { decimate(T1, T3, n, H, W, k.planes, st); t1 = T3; }
      run_conv(bb, bb->convs[k.c2], t1, k.downsample ? xs : nullptr, k.downsample ? nullptr : X, Y, n, Ho, Wo, 1, st);
    }
    float* tmp = X; X = Y; Y = tmp;
    H = Ho; W = Wo; cx = k.cout;
  }
  for (int d = 0; d < 3 && nb + d < stage; ++d) {
    const BbConv& c = bb->convs[bb->deconv[d]];
    run_conv(bb, c, X, nullptr, nullptr, Y, n, H, W, 1, st);                           // 3x3 + pixel shuffle: [H][W] -> [2H][2W]
    float* tmp = X; X = Y; Y = tmp;
    H *= 2; W *= 2; cx = c.cout;
  }
  if (stage == last) {
    run_conv(bb, bb->convs[bb->final_conv], X, nullptr, nullptr, nullptr, n, H, W, 0, st, d_out, bb->num_joints);
  } else {                                                 // a tap for the parity tests: split NHWC -> fp32 NHWC -> fp32 NCHW
    fvp_launch_unsplit(X, Y, (size_t)n * H * W * cx, st);
    fvp_launch_nhwc_to_nchw(Y, d_out, n, H * W, cx, cx, st);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bb_fail(bb, FVP_E_CUDA, "CUDA error: %s", cudaGetErrorString(e));
  if (bb->launch_error) { bb->launch_error = 0; return bb_fail(bb, FVP_E_CUDA, "a convolution launch could not be prepared"); }
  return FVP_OK;
}

// replaces ResNet.forward (resnet.py:188-201): [n][3][h][w] normalised images -> heat maps [n][J][h/4][w/4]
int fvp_backbone_forward(fvp_backbone* bb, const float* d_images, int n, int h, int w, float* d_heatmaps, uintptr_t stream) {
  if (!bb) return FVP_E_INVALID;
  return fvp_backbone_forward_slice(bb, d_images, n, h, w, (int)bb->blocks.size() + 4, d_heatmaps, stream);
}

}  // extern "C"
