// fvp_conv.cu - fp32 convolutions of the CenterNet / P2PNet trunks (cnns_2d.py:12-178) as NHWC
// implicit GEMMs on the CUDA cores, BatchNorm folded, with fused epilogues:
//   * bias (+BN) , residual add before or after ReLU, ReLU
//   * a second K segment for the 1x1 skip_con of a Res2DBlock (cnns_2d.py:38-47): conv3(y) + conv1(x)
//     accumulate into one tile, so a residual block is two launches
//   * ConvTranspose2d(k=2,s=2) (cnns_2d.py:58-71) = 1x1 conv to 4*Co channels + pixel-shuffle store
//   * planar (NCHW) store for the final layers the reference-layout consumers read
// This is the exact-fp32 path (summation order differs from cuDNN/oneDNN, nothing else).
#include <cuda_fp16.h>

#include "fvp_kernels.h"

// range guard twin of fvp_conv_tc.cu's g_tc_status: this kernel's outputs may feed a layer of the fp16 hi/lo engine
__device__ int* g_conv_status = nullptr;

namespace {

constexpr int TH = 8, TW = 8;          // output pixels per CTA
constexpr int CI_CHUNK = 16;           // input channels staged per pass
constexpr int NTHREADS = 128;

template <int CO_T>
struct ConvCfg {
  static constexpr int CO_GROUPS = CO_T / 4;                       // float4 groups along Cout
  static constexpr int PX_GROUPS = NTHREADS / CO_GROUPS;           // pixel groups
  static constexpr int PX = TH * TW / PX_GROUPS;                   // pixels per thread (4 or 2), consecutive in a row
};

__host__ __device__ inline int halo_stride(int k) {
  int hs = TW + k - 1;
  if ((hs & 7) == 0) hs += 2;      // keep the 4 pixel groups of a warp on distinct banks
  return hs;
}

// smem: A [CI_CHUNK/4][halo_h*hs] float4 , B [taps][CI_CHUNK][CO_T] float
template <int CO_T>
__global__ void __launch_bounds__(NTHREADS) k_conv_nhwc(FvpConvArgs a) {
  using C = ConvCfg<CO_T>;
  extern __shared__ float4 smem4[];
  const int img = blockIdx.z;
  if (a.valid && !a.valid[img]) return;
  const int tiles_x = (a.W + TW - 1) / TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  const int co0 = blockIdx.y * CO_T;

  const int tid = threadIdx.x;
  const int cog = tid % C::CO_GROUPS;             // my float4 of output channels
  const int pxg = tid / C::CO_GROUPS;             // my pixel group
  const int gpr = TW / C::PX;                     // groups per tile row
  const int py = pxg / gpr, px0 = (pxg - py * gpr) * C::PX;

  float4 acc[C::PX];
#pragma unroll
  for (int p = 0; p < C::PX; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);

  int wrow = 0;                                   // running row offset into the packed weights
  for (int phase = 0; phase < 2; ++phase) {
    const float* src = phase == 0 ? a.in : a.in2;
    if (src == nullptr) break;
    const int K = phase == 0 ? a.ksize : 1;
    const int Cin = phase == 0 ? a.Cin : a.Cin2;
    const int CinP = (Cin + CI_CHUNK - 1) / CI_CHUNK * CI_CHUNK;
    const int pad = (K - 1) / 2;
    const int hs = halo_stride(K), hh = TH + K - 1;
    const int npix = hs * hh;
    float4* sA = smem4;                           // [CI_CHUNK/4][npix]
    float* sB = (float*)(smem4 + (CI_CHUNK / 4) * npix);   // [K*K][CI_CHUNK][CO_T]
    const float* img_in = src + (size_t)img * a.H * a.W * Cin;

    for (int c0 = 0; c0 < CinP; c0 += CI_CHUNK) {
      __syncthreads();                            // previous chunk fully consumed
      // ---- stage the input halo chunk (zero padded) ----
      for (int i = tid; i < npix * (CI_CHUNK / 4); i += NTHREADS) {
        const int q = i % (CI_CHUNK / 4), pix = i / (CI_CHUNK / 4);
        const int hy = pix / hs, hx = pix - hy * hs;
        const int gy = y0 + hy - pad, gx = x0 + hx - pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = c0 + q * 4;
        if (hx < TW + K - 1 && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && c < Cin)
          v = __ldg((const float4*)(img_in + ((size_t)gy * a.W + gx) * Cin + c));
        sA[q * npix + pix] = v;
      }
      // ---- stage the weight chunk: rows (tap, c0..c0+15), cols co0..co0+CO_T ----
      for (int i = tid; i < K * K * CI_CHUNK * (CO_T / 4); i += NTHREADS) {
        const int cq = i % (CO_T / 4), r = i / (CO_T / 4);
        const int ci = r % CI_CHUNK, tap = r / CI_CHUNK;
        const int co = co0 + cq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co < a.CoutP)
          v = __ldg((const float4*)(a.w + (size_t)(wrow + tap * CinP + c0 + ci) * a.CoutP + co));
        ((float4*)sB)[r * (CO_T / 4) + cq] = v;
      }
      __syncthreads();
      // ---- multiply ----
      for (int tap = 0; tap < K * K; ++tap) {
        const int dy = tap / K, dx = tap - dy * K;
        const float4* arow = sA + (py + dy) * hs + px0 + dx;
        const float4* brow = (const float4*)sB + (size_t)tap * CI_CHUNK * (CO_T / 4) + cog;
#pragma unroll
        for (int q = 0; q < CI_CHUNK / 4; ++q) {
          float4 av[C::PX];
#pragma unroll
          for (int p = 0; p < C::PX; ++p) av[p] = arow[q * npix + p];
          const float4 b0 = brow[(q * 4 + 0) * (CO_T / 4)];
          const float4 b1 = brow[(q * 4 + 1) * (CO_T / 4)];
          const float4 b2 = brow[(q * 4 + 2) * (CO_T / 4)];
          const float4 b3 = brow[(q * 4 + 3) * (CO_T / 4)];
#pragma unroll
          for (int p = 0; p < C::PX; ++p) {
            acc[p].x = fmaf(av[p].w, b3.x, fmaf(av[p].z, b2.x, fmaf(av[p].y, b1.x, fmaf(av[p].x, b0.x, acc[p].x))));
            acc[p].y = fmaf(av[p].w, b3.y, fmaf(av[p].z, b2.y, fmaf(av[p].y, b1.y, fmaf(av[p].x, b0.y, acc[p].y))));
            acc[p].z = fmaf(av[p].w, b3.z, fmaf(av[p].z, b2.z, fmaf(av[p].y, b1.z, fmaf(av[p].x, b0.z, acc[p].z))));
            acc[p].w = fmaf(av[p].w, b3.w, fmaf(av[p].z, b2.w, fmaf(av[p].y, b1.w, fmaf(av[p].x, b0.w, acc[p].w))));
          }
        }
      }
    }
    wrow += K * K * CinP;
  }

  // ---- epilogue ----
  const int co = co0 + cog * 4;
  if (co >= a.CoutP) return;
  const float4 bias = __ldg((const float4*)(a.bias + co));
  const int oy = y0 + py;
  if (oy >= a.H) return;
#pragma unroll
  for (int p = 0; p < C::PX; ++p) {
    const int ox = x0 + px0 + p;
    if (ox >= a.W) continue;
    float4 v = make_float4(acc[p].x + bias.x, acc[p].y + bias.y, acc[p].z + bias.z, acc[p].w + bias.w);
    int Y = oy, X = ox, Ho = a.H, Wo = a.W, ch = co;
    if (a.upsample) {                             // co' = q*Co + c, q = dy*2+dx (ConvTranspose k2 s2)
      const int Co = a.CoutP >> 2;
      const int q = co / Co;
      ch = co - q * Co;
      Y = 2 * oy + (q >> 1);
      X = 2 * ox + (q & 1);
      Ho = 2 * a.H;
      Wo = 2 * a.W;
    }
    const size_t opix = ((size_t)img * Ho + Y) * Wo + X;
    if (a.res_mode == 1) {
      const float4 r = __ldg((const float4*)(a.res + opix * a.CoutS + ch));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (a.res_mode == 2) {
      const float4 r = __ldg((const float4*)(a.res + opix * a.CoutS + ch));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (!(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) < 65504.0f) && g_conv_status) *g_conv_status = 1;
    if (!a.nchw) {
      *(float4*)(a.out + opix * a.CoutS + ch) = v;
    } else {
      const size_t plane = (size_t)Ho * Wo;
      float* o = a.out + (size_t)img * a.CoutReal * plane + (size_t)Y * Wo + X;
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (ch + e < a.CoutReal) o[(size_t)(ch + e) * plane] = vv[e];
    }
  }
}

size_t conv_smem_bytes(int k, int co_t) {
  const int hs = halo_stride(k), hh = TH + k - 1;
  return (size_t)(CI_CHUNK / 4) * hs * hh * 16 + (size_t)k * k * CI_CHUNK * co_t * 4;
}

__global__ void __launch_bounds__(256) k_maxpool2(const float4* __restrict__ in, float4* __restrict__ out, int H,
                                                  int W, int C4, const int* __restrict__ valid) {
  const int img = blockIdx.y;
  if (valid && !valid[img]) return;
  const int Ho = H >> 1, Wo = W >> 1;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Ho * Wo * C4) return;
  const int c = i % C4, p = i / C4;
  const int oy = p / Wo, ox = p - oy * Wo;
  const float4* s = in + ((size_t)img * H * W + (size_t)(2 * oy) * W + 2 * ox) * C4 + c;
  const float4 a = s[0], b = s[C4], d = s[(size_t)W * C4], e = s[(size_t)W * C4 + C4];
  out[(size_t)img * Ho * Wo * C4 + i] =
      make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                  fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
}

// 2x2 max-pool on a split tensor: 8 channels (one uint4 of hi halves + one of lo halves) per thread.  The winner is chosen
// on the reconstructed values hi + lo * 2^-11 and its (hi, lo) pair is copied, so pooling a split tensor is exact.
__global__ void __launch_bounds__(256) k_maxpool2_split(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int H, int W,
                                                        int C8, const int* __restrict__ valid) {
  const int img = blockIdx.y;
  if (valid && !valid[img]) return;
  const int Ho = H >> 1, Wo = W >> 1;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Ho * Wo * C8) return;
  const int c = i % C8, p = i / C8;
  const int oy = p / Wo, ox = p - oy * Wo;
  const size_t plane_in = (size_t)n * H * W * C8, plane_out = (size_t)n * Ho * Wo * C8;
  const size_t base = ((size_t)img * H * W + (size_t)(2 * oy) * W + 2 * ox) * C8 + c;
  const size_t offs[4] = {0, (size_t)C8, (size_t)W * C8, (size_t)W * C8 + C8};
  uint4 bh = in[base], bl = in[plane_in + base];
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    const uint4 h = in[base + offs[q]], l = in[plane_in + base + offs[q]];
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(&h);
    const uint32_t* lw = reinterpret_cast<const uint32_t*>(&l);
    uint32_t* bhw = reinterpret_cast<uint32_t*>(&bh);
    uint32_t* blw = reinterpret_cast<uint32_t*>(&bl);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 xh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e])), xl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
      const float2 yh = __half22float2(*reinterpret_cast<const __half2*>(&bhw[e])), yl = __half22float2(*reinterpret_cast<const __half2*>(&blw[e]));
      const bool t0 = fmaf(xl.x, 1.0f / 2048.0f, xh.x) > fmaf(yl.x, 1.0f / 2048.0f, yh.x);
      const bool t1 = fmaf(xl.y, 1.0f / 2048.0f, xh.y) > fmaf(yl.y, 1.0f / 2048.0f, yh.y);
      const uint32_t m = (t0 ? 0x0000ffffu : 0u) | (t1 ? 0xffff0000u : 0u);
      bhw[e] = (hw[e] & m) | (bhw[e] & ~m);
      blw[e] = (lw[e] & m) | (blw[e] & ~m);
    }
  }
  const size_t o = (size_t)img * Ho * Wo * C8 + i;
  out[o] = bh;
  out[plane_out + o] = bl;
}

// test / transition helpers: fp32 <-> split (hi, scaled lo) element-wise over `count` values
__global__ void k_split(const float* __restrict__ in, __half* __restrict__ out, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float x = in[i];
  const __half h = __float2half_rn(x);
  out[i] = h;
  out[count + i] = __float2half_rn((x - __half2float(h)) * 2048.0f);
}
__global__ void k_unsplit(const __half* __restrict__ in, float* __restrict__ out, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[i] = fmaf(__half2float(in[count + i]), 1.0f / 2048.0f, __half2float(in[i]));
}

}  // namespace

void fvp_launch_split(const float* in, void* out, size_t count, cudaStream_t st) {
  k_split<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(in, (__half*)out, count);
}
void fvp_launch_unsplit(const void* in, float* out, size_t count, cudaStream_t st) {
  k_unsplit<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const __half*)in, out, count);
}

cudaError_t fvp_conv_set_status_ptr(int* d_status) { return cudaMemcpyToSymbol(g_conv_status, &d_status, sizeof(d_status)); }

// per-device: the 7x7 variants need more than the default 48 KB of dynamic shared memory (called by fvp_create)
cudaError_t fvp_conv_init_device() {
  cudaError_t e = cudaFuncSetAttribute(k_conv_nhwc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_bytes(7, 16));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_conv_nhwc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_bytes(7, 32));
}

void fvp_launch_conv(const FvpConvArgs& a, cudaStream_t st) {
  const int tiles = fvp_cdiv(a.H, TH) * fvp_cdiv(a.W, TW);
  if (a.CoutP <= 16) {
    dim3 grid(tiles, 1, a.n);
    k_conv_nhwc<16><<<grid, NTHREADS, conv_smem_bytes(a.ksize, 16), st>>>(a);
  } else {
    dim3 grid(tiles, fvp_cdiv(a.CoutP, 32), a.n);
    k_conv_nhwc<32><<<grid, NTHREADS, conv_smem_bytes(a.ksize, 32), st>>>(a);
  }
}

void fvp_launch_maxpool2(const float* in, float* out, int n, int H, int W, int C, const int* valid, cudaStream_t st) {
  dim3 grid(fvp_cdiv((H / 2) * (W / 2) * (C / 4), 256), n);
  k_maxpool2<<<grid, 256, 0, st>>>((const float4*)in, (float4*)out, H, W, C / 4, valid);
}
void fvp_launch_maxpool2_split(const void* in, void* out, int n, int H, int W, int C, const int* valid, cudaStream_t st) {
  dim3 grid(fvp_cdiv((H / 2) * (W / 2) * (C / 8), 256), n);
  k_maxpool2_split<<<grid, 256, 0, st>>>((const uint4*)in, (uint4*)out, n, H, W, C / 8, valid);
}

// ------------------------------------------------------------------------------------------------
// trunk program: front_layers + EncoderDecorder (+ heads), cnns_2d.py:94-135,173-178
// ------------------------------------------------------------------------------------------------
namespace {
struct TrunkRun {            // what every conv of one trunk enqueue shares (no process-global launcher state)
  const FvpLaunchEnv& env;
  int n;
  const int* valid;
  int* launches;
  cudaStream_t st;
};
}  // namespace
static void conv(const TrunkRun& r, const FvpConvW& w, const float* in, int H, int W, const float* in2, float* out, int couts,
                 const float* res, int res_mode, int relu, int upsample, int nchw = 0, int cout_real = 0, int fmt = 0) {
  const int n = r.n;
  const int* valid = r.valid;
  int* launches = r.launches;
  cudaStream_t st = r.st;
  const int tc = r.env.conv_mode;
  FvpConvArgs a;
  a.in = in; a.H = H; a.W = W; a.Cin = w.cin;
  a.in2 = in2; a.Cin2 = w.cin2;
  a.w = w.w; a.bias = w.b;
  a.out = out; a.CoutP = w.coutp; a.CoutS = couts; a.CoutReal = cout_real;
  a.res = res; a.res_mode = res_mode; a.relu = relu; a.ksize = w.k; a.upsample = upsample;
  a.nchw = nchw; a.n = n; a.valid = valid;
  a.fmt = fmt;
  if (fmt & FVP_FMT_IN_SPLIT) {                     // TMA-fed layer of a split-activation trunk (fp16 engine by construction)
    const float* const c16s[3] = {w.wtc16_c16, nullptr, nullptr};
    if (w.wtc16_c16 && w.cin <= 16 && w.k != 7) fvp_launch_conv_tc(a, c16s, 2, r.env, st);
    else fvp_launch_conv_tc(a, w.wtc16, 1, r.env, st);
    if (launches) ++*launches;
    return;
  }
  // engine 2: fp16-split tcgen05 kernel (all layers); engine 1: 3xTF32 tcgen05 kernel, where the 7x7 front conv (49 taps
  // of half-empty 32-channel K-blocks, measured 6.6 vs 11.4 TMAC/s) stays on the CUDA-core kernel; engine 0: CUDA cores.
  // A layer whose BN-folded weights leave the fp16 range has no fp16 image (fvp_params.cu: stash) and drops to the 3xTF32
  // engine (fp32 exponent range) or, for a 7x7, to the CUDA cores.
  const float* const c16[3] = {w.wtc16_c16, nullptr, nullptr};
  if (tc == 2 && w.wtc16_c16) fvp_launch_conv_tc(a, c16, 2, r.env, st);             // 16-channel K-blocks (<= 16 channels, 7x7)
  else if (tc == 2 && w.wtc16[0] && w.k != 7) fvp_launch_conv_tc(a, w.wtc16, 1, r.env, st);
  else if (tc >= 1 && w.wtc[0] && w.k != 7) fvp_launch_conv_tc(a, w.wtc, 0, r.env, st);
  else fvp_launch_conv(a, st);
  if (launches) ++*launches;
}

void fvp_run_trunk2d(const FvpTrunkW& t, const float* d_in, int cin, int n, int H, int W, float* const buf[6],
                     const int* valid, bool center_heads, float* d_out, int out_real, int* launches, cudaStream_t st,
                     const FvpLaunchEnv& env) {
  (void)cin;
  const TrunkRun r{env, n, valid, launches, st};
  float *B0 = buf[0], *B1 = buf[1], *B2 = buf[2], *B3 = buf[3], *B4 = buf[4], *B5 = buf[5];
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  // Engine 2 with an fp16 image for every layer (the normal case): activations travel between layers as SPLIT tensors -
  // every epilogue stores hi / scaled-lo fp16 planes once, every consumer fetches its halos with TMA tensor loads (no
  // loader warps, no conversions, zero padding = TMA out-of-bounds fill).  Only the first layer reads fp32 (the
  // back-projected planes) through the legacy loaders and only the last one writes fp32.  Scratch units hold either form
  // (2 x fp16 = 1 x fp32 per element).
  const FvpConvW* all[21] = {&t.front, &t.r1a, &t.r1b, &t.s1a, &t.s1b, &t.e1a, &t.e1b, &t.s2a, &t.s2b, &t.e2a, &t.e2b,
                             &t.ma, &t.mb, &t.d2a, &t.d2b, &t.up2, &t.d1a, &t.d1b, &t.up1, &t.head_b, center_heads ? &t.head_a : &t.head_b};
  bool split = env.conv_mode == 2 && env.split_activations && t.front.wtc16_c16 != nullptr;
  for (int i = 1; i < 21 && split; ++i) split = all[i]->wtc16[0] != nullptr;
  if (split) {
    const int S = FVP_FMT_IN_SPLIT | FVP_FMT_OUT_SPLIT, SR = S | FVP_FMT_RES_SPLIT;
    conv(r, t.front, d_in, H, W, nullptr, B0, 16, nullptr, 0, 1, 0, 0, 0, FVP_FMT_OUT_SPLIT);   // fp32 in (legacy loaders), split out
    conv(r, t.r1a, B0, H, W, nullptr, B1, 32, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.r1b, B1, H, W, B0, B2, 32, nullptr, 0, 1, 0, 0, 0, S);                             // f1 (conv skip fused)
    conv(r, t.s1a, B2, H, W, nullptr, B0, 32, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.s1b, B0, H, W, nullptr, B3, 32, B2, 1, 1, 0, 0, 0, SR);                            // skip1
    fvp_launch_maxpool2_split(B2, B0, n, H, W, 32, valid, st); if (launches) ++*launches;
    conv(r, t.e1a, B0, H2, W2, nullptr, B1, 64, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.e1b, B1, H2, W2, B0, B4, 64, nullptr, 0, 1, 0, 0, 0, S);                           // e1
    conv(r, t.s2a, B4, H2, W2, nullptr, B0, 64, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.s2b, B0, H2, W2, nullptr, B5, 64, B4, 1, 1, 0, 0, 0, SR);                          // skip2
    fvp_launch_maxpool2_split(B4, B0, n, H2, W2, 64, valid, st); if (launches) ++*launches;
    conv(r, t.e2a, B0, H4, W4, nullptr, B1, 128, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.e2b, B1, H4, W4, B0, B2, 128, nullptr, 0, 1, 0, 0, 0, S);                          // e2
    conv(r, t.ma, B2, H4, W4, nullptr, B0, 128, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.mb, B0, H4, W4, nullptr, B1, 128, B2, 1, 1, 0, 0, 0, SR);                          // m
    conv(r, t.d2a, B1, H4, W4, nullptr, B0, 128, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.d2b, B0, H4, W4, nullptr, B2, 128, B1, 1, 1, 0, 0, 0, SR);                         // d2
    conv(r, t.up2, B2, H4, W4, nullptr, B0, 64, B5, 2, 1, 1, 0, 0, SR);                          // u2 @ half
    conv(r, t.d1a, B0, H2, W2, nullptr, B1, 64, nullptr, 0, 1, 0, 0, 0, S);
    conv(r, t.d1b, B1, H2, W2, nullptr, B2, 64, B0, 1, 1, 0, 0, 0, SR);                          // d1
    conv(r, t.up1, B2, H2, W2, nullptr, B0, 32, B3, 2, 1, 1, 0, 0, SR);                          // u1 @ full
    if (center_heads) {
      conv(r, t.head_a, B0, H, W, nullptr, B1, 64, nullptr, 0, 1, 0, 0, 0, S);                   // both 3x3 heads + ReLU
      conv(r, t.head_b, B1, H, W, nullptr, d_out, 0, nullptr, 0, 0, 0, 1, out_real, FVP_FMT_IN_SPLIT);
    } else {
      conv(r, t.head_b, B0, H, W, nullptr, d_out, 0, nullptr, 0, 0, 0, 1, out_real, FVP_FMT_IN_SPLIT);
    }
    return;
  }
  // front_layers
  conv(r, t.front, d_in, H, W, nullptr, B0, 16, nullptr, 0, 1, 0);          // t16
  conv(r, t.r1a, B0, H, W, nullptr, B1, 32, nullptr, 0, 1, 0);
  conv(r, t.r1b, B1, H, W, B0, B2, 32, nullptr, 0, 1, 0);                     // f1 (conv skip fused)
  // skip_res1 @ full
  conv(r, t.s1a, B2, H, W, nullptr, B0, 32, nullptr, 0, 1, 0);
  conv(r, t.s1b, B0, H, W, nullptr, B3, 32, B2, 1, 1, 0);                     // skip1
  // encoder_res1 @ half
  fvp_launch_maxpool2(B2, B0, n, H, W, 32, valid, st); if (launches) ++*launches;
  conv(r, t.e1a, B0, H2, W2, nullptr, B1, 64, nullptr, 0, 1, 0);
  conv(r, t.e1b, B1, H2, W2, B0, B4, 64, nullptr, 0, 1, 0);                   // e1
  // skip_res2 @ half
  conv(r, t.s2a, B4, H2, W2, nullptr, B0, 64, nullptr, 0, 1, 0);
  conv(r, t.s2b, B0, H2, W2, nullptr, B5, 64, B4, 1, 1, 0);                   // skip2
  // encoder_res2 @ quarter
  fvp_launch_maxpool2(B4, B0, n, H2, W2, 64, valid, st); if (launches) ++*launches;
  conv(r, t.e2a, B0, H4, W4, nullptr, B1, 128, nullptr, 0, 1, 0);
  conv(r, t.e2b, B1, H4, W4, B0, B2, 128, nullptr, 0, 1, 0);                  // e2
  // mid_res, decoder_res2
  conv(r, t.ma, B2, H4, W4, nullptr, B0, 128, nullptr, 0, 1, 0);
  conv(r, t.mb, B0, H4, W4, nullptr, B1, 128, B2, 1, 1, 0);                   // m
  conv(r, t.d2a, B1, H4, W4, nullptr, B0, 128, nullptr, 0, 1, 0);
  conv(r, t.d2b, B0, H4, W4, nullptr, B2, 128, B1, 1, 1, 0);                  // d2
  // decoder_upsample2 (+ skip2 after ReLU), decoder_res1
  conv(r, t.up2, B2, H4, W4, nullptr, B0, 64, B5, 2, 1, 1);                   // u2 @ half
  conv(r, t.d1a, B0, H2, W2, nullptr, B1, 64, nullptr, 0, 1, 0);
  conv(r, t.d1b, B1, H2, W2, nullptr, B2, 64, B0, 1, 1, 0);                   // d1
  // decoder_upsample1 (+ skip1)
  conv(r, t.up1, B2, H2, W2, nullptr, B0, 32, B3, 2, 1, 1);                   // u1 @ full
  if (center_heads) {
    conv(r, t.head_a, B0, H, W, nullptr, B1, 64, nullptr, 0, 1, 0);           // both 3x3 heads + ReLU
    conv(r, t.head_b, B1, H, W, nullptr, d_out, 0, nullptr, 0, 0, 0, 1, out_real);
  } else {
    conv(r, t.head_b, B0, H, W, nullptr, d_out, 0, nullptr, 0, 0, 0, 1, out_real);
  }
}
