// fvp_params.cu - the reference's state_dict as a parameter table, eval-mode BatchNorm folding and
// repacking for the kernels.  Table order == nn.Module.state_dict() order of FasterVoxelPoseNet
// (485 entries for J=15; SURVEY.md section 5 "checkpoint"); tests/test_netspec.py checks it against
// the Python enumeration (fvp/netspec.py), which itself was checked against the reference.
#include <cmath>
#include <cstring>

#include <cuda_fp16.h>

#include "fvp_ctx.h"
#include <cstdlib>

namespace {

void add_res(std::vector<FvpLayer>& L, const std::string& p, int cin, int cout, int nd) {
  L.push_back({p + ".res_branch.0", p + ".res_branch.1", cin, cout, 3, nd, false});
  L.push_back({p + ".res_branch.3", p + ".res_branch.4", cout, cout, 3, nd, false});
  if (cin != cout) L.push_back({p + ".skip_con.0", p + ".skip_con.1", cin, cout, 1, nd, false});
}

// front_layers + EncoderDecorder in registration order (cnns_2d.py:78-92,150-155)
void add_trunk(std::vector<FvpLayer>& L, const std::string& p, int cin, int nd) {
  const std::string ed = p + ".encoder_decoder";
  L.push_back({p + ".front_layers.0.block.0", p + ".front_layers.0.block.1", cin, 16, 7, nd, false});
  add_res(L, p + ".front_layers.1", 16, 32, nd);
  add_res(L, ed + ".encoder_res1", 32, 64, nd);
  add_res(L, ed + ".encoder_res2", 64, 128, nd);
  add_res(L, ed + ".mid_res", 128, 128, nd);
  add_res(L, ed + ".decoder_res2", 128, 128, nd);
  L.push_back({ed + ".decoder_upsample2.block.0", ed + ".decoder_upsample2.block.1", 128, 64, 2, nd, true});
  add_res(L, ed + ".decoder_res1", 64, 64, nd);
  L.push_back({ed + ".decoder_upsample1.block.0", ed + ".decoder_upsample1.block.1", 64, 32, 2, nd, true});
  add_res(L, ed + ".skip_res1", 32, 32, nd);
  add_res(L, ed + ".skip_res2", 64, 64, nd);
}

int64_t ipow(int b, int e) {
  int64_t r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

}  // namespace

void fvp_build_param_table(fvp_ctx* ctx) {
  const int J = ctx->cfg.num_joints;
  std::vector<FvpLayer>& L = ctx->layers;
  L.clear();
  add_trunk(L, "pose_net.center_net", J, 2);
  L.push_back({"pose_net.center_net.output_hm.0", "", 32, 32, 3, 2, false});
  L.push_back({"pose_net.center_net.output_hm.2", "", 32, 1, 1, 2, false});
  L.push_back({"pose_net.center_net.output_size.0", "", 32, 32, 3, 2, false});
  L.push_back({"pose_net.center_net.output_size.2", "", 32, 2, 1, 2, false});
  add_trunk(L, "pose_net.c2c_net", J, 1);
  L.push_back({"pose_net.c2c_net.output_hm", "", 32, 1, 1, 1, false});
  add_trunk(L, "joint_net.conv_net", J, 2);
  L.push_back({"joint_net.conv_net.output_layer", "", 32, J, 1, 2, false});
  const int F = ctx->cfg.feat_channels, Hd = ctx->cfg.hidden_channels;
  L.push_back({"joint_net.weight_net.heatmap_feature_net.0", "joint_net.weight_net.heatmap_feature_net.1", 1, F, 3, 2, false});
  L.push_back({"joint_net.weight_net.output.0", "", F, Hd, 1, 0, false});
  L.push_back({"joint_net.weight_net.output.2", "", Hd, 1, 1, 0, false});

  ctx->params.clear();
  ctx->param_index.clear();
  auto add = [&](const std::string& name, int64_t numel, bool is_int) {
    FvpParam p;
    p.name = name;
    p.numel = numel;
    p.is_int = is_int;
    p.set = is_int;  // num_batches_tracked is optional
    if (!is_int) p.data.assign((size_t)numel, 0.f);
    ctx->param_index[name] = (int)ctx->params.size();
    ctx->params.push_back(std::move(p));
  };
  for (const FvpLayer& l : L) {
    const int64_t kk = l.ndim ? ipow(l.k, l.ndim) : 1;
    add(l.key + ".weight", (int64_t)l.cin * l.cout * kk, false);
    add(l.key + ".bias", l.cout, false);
    if (!l.bn.empty()) {
      add(l.bn + ".weight", l.cout, false);
      add(l.bn + ".bias", l.cout, false);
      add(l.bn + ".running_mean", l.cout, false);
      add(l.bn + ".running_var", l.cout, false);
      add(l.bn + ".num_batches_tracked", 1, true);
    }
  }
}

namespace {

const FvpLayer* find_layer(const fvp_ctx* ctx, const std::string& key) {
  for (const FvpLayer& l : ctx->layers)
    if (l.key == key) return &l;
  return nullptr;
}
const std::vector<float>& P(const fvp_ctx* ctx, const std::string& name) {
  return ctx->params[ctx->param_index.at(name)].data;
}

// BN scale / shift of a layer's output channels (identity when no BN): y = s*conv + t,
// s = gamma / sqrt(var + 1e-5), t = (bias - mean) * s + beta        (F.batch_norm, eval)
void bn_affine(const fvp_ctx* ctx, const FvpLayer& l, std::vector<double>& s, std::vector<double>& t) {
  const std::vector<float>& b = P(ctx, l.key + ".bias");
  s.assign(l.cout, 1.0);
  t.assign(l.cout, 0.0);
  for (int c = 0; c < l.cout; ++c) t[c] = b[c];
  if (l.bn.empty()) return;
  const std::vector<float>&g = P(ctx, l.bn + ".weight"), &be = P(ctx, l.bn + ".bias"),
                    &m = P(ctx, l.bn + ".running_mean"), &v = P(ctx, l.bn + ".running_var");
  for (int c = 0; c < l.cout; ++c) {
    s[c] = (double)g[c] / std::sqrt((double)v[c] + 1e-5);
    t[c] = ((double)b[c] - (double)m[c]) * s[c] + (double)be[c];
  }
}

struct Packed {
  std::vector<float> w, b;
  int cin, cin2, coutp, k;
};

// regular conv (2-D or 1-D), optional fused 1x1 skip conv as extra K rows
Packed pack_conv(const fvp_ctx* ctx, const std::string& key, const std::string& skip_key = "") {
  const FvpLayer& l = *find_layer(ctx, key);
  const int nd = l.ndim, K = l.k, taps = (int)ipow(K, nd);
  const int cinP = fvp_round_up(l.cin, 16), coutP = fvp_round_up(l.cout, 4);
  const FvpLayer* sk = skip_key.empty() ? nullptr : find_layer(ctx, skip_key);
  const int cin2 = sk ? sk->cin : 0, cin2P = sk ? fvp_round_up(cin2, 16) : 0;
  Packed p;
  p.cin = l.cin; p.cin2 = cin2; p.coutp = coutP; p.k = K;
  p.w.assign((size_t)(taps * cinP + cin2P) * coutP, 0.f);
  p.b.assign(coutP, 0.f);
  std::vector<double> s, t;
  bn_affine(ctx, l, s, t);
  const std::vector<float>& w = P(ctx, key + ".weight");      // [co][ci][taps]
  for (int co = 0; co < l.cout; ++co)
    for (int ci = 0; ci < l.cin; ++ci)
      for (int tp = 0; tp < taps; ++tp)
        p.w[(size_t)(tp * cinP + ci) * coutP + co] = (float)((double)w[((size_t)co * l.cin + ci) * taps + tp] * s[co]);
  std::vector<double> bias(t);
  if (sk) {
    std::vector<double> s2, t2;
    bn_affine(ctx, *sk, s2, t2);
    const std::vector<float>& w2 = P(ctx, skip_key + ".weight");   // [co][ci][1]
    for (int co = 0; co < l.cout; ++co) {
      for (int ci = 0; ci < cin2; ++ci)
        p.w[(size_t)(taps * cinP + ci) * coutP + co] = (float)((double)w2[(size_t)co * cin2 + ci] * s2[co]);
      bias[co] += t2[co];
    }
  }
  for (int co = 0; co < l.cout; ++co) p.b[co] = (float)bias[co];
  return p;
}

// ConvTranspose(k=2,s=2): [ci][co][q] -> GEMM columns q*Co + co, q = dy*2+dx (2-D) or d (1-D)
Packed pack_convT(const fvp_ctx* ctx, const std::string& key) {
  const FvpLayer& l = *find_layer(ctx, key);
  const int Q = (int)ipow(2, l.ndim);
  const int cinP = fvp_round_up(l.cin, 16), coutP = Q * l.cout;
  Packed p;
  p.cin = l.cin; p.cin2 = 0; p.coutp = coutP; p.k = 1;
  p.w.assign((size_t)cinP * coutP, 0.f);
  p.b.assign(coutP, 0.f);
  std::vector<double> s, t;
  bn_affine(ctx, l, s, t);
  const std::vector<float>& w = P(ctx, key + ".weight");      // [ci][co][Q]
  for (int ci = 0; ci < l.cin; ++ci)
    for (int co = 0; co < l.cout; ++co)
      for (int q = 0; q < Q; ++q)
        p.w[(size_t)ci * coutP + q * l.cout + co] = (float)((double)w[((size_t)ci * l.cout + co) * Q + q] * s[co]);
  for (int q = 0; q < Q; ++q)
    for (int co = 0; co < l.cout; ++co) p.b[q * l.cout + co] = (float)t[co];
  return p;
}

struct Arena {
  std::vector<float> host;
  size_t put(const std::vector<float>& v) {
    const size_t off = (host.size() + 63) & ~(size_t)63;     // 256-byte aligned
    host.resize(off + v.size());
    std::memcpy(host.data() + off, v.data(), v.size() * sizeof(float));
    return off;
  }
};

struct PendingConv {
  size_t w_off, b_off, tc_off[3], t16_off[3], t16c_off;
  int cin, cin2, coutp, k;
};

// cvt.rna.tf32.f32 on the host: round to nearest (ties away), keep 10 mantissa bits
float tf32_rna(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

// Tiled tf32 hi/lo image of a packed conv for k_conv_tc: blocks [phase][32-channel K-block][tap][N-tile], each block =
// hi[n_tile rows][128 B] then lo[...], rows K-major with the 16-B chunks XOR-swizzled by (row & 7) (SWIZZLE_128B).
// (cin_act = channels of the activation tensor the kernel will see: JP for the first layer)
std::vector<float> pack_tc(const Packed& p, int cin_act, int narrow) {
  int n_tile, n_tiles;
  fvp_tc_geometry(p.coutp, narrow, &n_tile, &n_tiles);
  (void)cin_act;
  const int taps = p.k * p.k, cinP = fvp_round_up(p.cin, 16), cin2P = p.cin2 ? fvp_round_up(p.cin2, 16) : 0;
  std::vector<float> out;
  for (int ph = 0; ph < (p.cin2 ? 2 : 1); ++ph) {
    const int K2 = ph == 0 ? taps : 1, CP = ph == 0 ? cinP : cin2P;
    const int rowbase = ph == 0 ? 0 : taps * cinP;
    const int CP32 = fvp_round_up(CP, 32);
    for (int c0 = 0; c0 < CP32; c0 += 32)
      for (int tap = 0; tap < K2; ++tap)
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int part = 0; part < 2; ++part) {
            const size_t base = out.size();
            out.resize(base + (size_t)n_tile * 32, 0.f);
            for (int n = 0; n < n_tile; ++n)
              for (int cc = 0; cc < 32; ++cc) {
                const int col = nt * n_tile + n, ci = c0 + cc;
                float w = 0.f;
                if (col < p.coutp && ci < CP) w = p.w[(size_t)(rowbase + tap * CP + ci) * p.coutp + col];
                const float hi = tf32_rna(w);
                const int chunk = (cc >> 2) ^ (n & 7);
                out[base + (size_t)n * 32 + chunk * 4 + (cc & 3)] = part == 0 ? hi : tf32_rna(w - hi);
              }
          }
  }
  return out;
}

// fp16 variant of pack_tc: rows of 64 B (32 halves), chunks XOR-swizzled by (row >> 1) & 3 (SWIZZLE_64B on absolute
// addresses, blocks are 512-B aligned); hi = fp16(w), lo = fp16((w - hi) * 2^11).  Returned as raw bytes in floats.
std::vector<float> pack_tc16(const Packed& p, int narrow, int cb = 32) {
  int n_tile, n_tiles;
  fvp_tc_geometry(p.coutp, narrow, &n_tile, &n_tiles);
  const int chunks = cb / 8, ph_shift = cb == 32 ? 1 : 2;        // rows of cb halves: SWIZZLE_64B / SWIZZLE_32B
  const int taps = p.k * p.k, cinP = fvp_round_up(p.cin, 16), cin2P = p.cin2 ? fvp_round_up(p.cin2, 16) : 0;
  std::vector<uint16_t> out;
  for (int ph = 0; ph < (p.cin2 ? 2 : 1); ++ph) {
    const int K2 = ph == 0 ? taps : 1, CP = ph == 0 ? cinP : cin2P;
    const int rowbase = ph == 0 ? 0 : taps * cinP;
    const int CPB = fvp_round_up(CP, cb);
    for (int c0 = 0; c0 < CPB; c0 += cb)
      for (int tap = 0; tap < K2; ++tap)
        for (int nt = 0; nt < n_tiles; ++nt)
          for (int part = 0; part < 2; ++part) {
            const size_t base = out.size();
            out.resize(base + (size_t)n_tile * cb, 0);
            for (int n = 0; n < n_tile; ++n)
              for (int cc = 0; cc < cb; ++cc) {
                const int col = nt * n_tile + n, ci = c0 + cc;
                float w = 0.f;
                if (col < p.coutp && ci < CP) w = p.w[(size_t)(rowbase + tap * CP + ci) * p.coutp + col];
                const __half hi = __float2half_rn(w);
                const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.0f);
                const int chunk = (cc >> 3) ^ ((n >> ph_shift) & (chunks - 1));
                out[base + (size_t)n * cb + chunk * 8 + (cc & 7)] = __half_as_ushort(part == 0 ? hi : lo);
              }
          }
  }
  std::vector<float> f((out.size() + 1) / 2, 0.f);
  std::memcpy(f.data(), out.data(), out.size() * sizeof(uint16_t));
  return f;
}

// Largest magnitude the fp16 hi/lo engine may be handed as a weight: fp16(w) must be finite (|w| < 65520 rounds to <= 65504).
// Real checkpoints can exceed it after BN folding (small running_var, large gamma); such a layer gets no fp16 image and
// runs on the 3xTF32 engine (fp32 exponent range) instead - see conv() in fvp_conv.cu.
constexpr float FVP_FP16_WEIGHT_LIMIT = 65504.0f;
bool fits_fp16(const Packed& p) {
  float m = 0.f;
  for (float w : p.w) m = std::fmax(m, std::fabs(w));                 // NaN-safe: fmax ignores NaN, caught below
  for (float w : p.w) if (!(std::fabs(w) < INFINITY)) return false;
  return m < FVP_FP16_WEIGHT_LIMIT;
}

PendingConv stash(Arena& A, const Packed& p, bool tc = false, int cin_act = 0) {
  PendingConv pc;
  pc.w_off = A.put(p.w);
  pc.b_off = A.put(p.b);
  const int npad = fvp_round_up(p.coutp, 16);
  const bool f16_ok = fits_fp16(p);
  for (int v = 0; v < 3; ++v) {                      // N-tile caps 128 / 32 / 64 (only where they differ from the wide image)
    const bool need = tc && (v == 0 || (v == 1 && npad > 32) || (v == 2 && npad > 64));
    pc.tc_off[v] = need ? A.put(pack_tc(p, cin_act ? cin_act : p.cin, v)) : (size_t)-1;
    pc.t16_off[v] = need && f16_ok ? A.put(pack_tc16(p, v)) : (size_t)-1;
  }
  // layers whose (activation) input has <= 16 channels also get a 16-channel K-block image (32-B rows); so does the 7x7
  // front conv with up to 32 input channels (J = 17 -> JP = 20: Campus / Shelf), as two 16-channel K-blocks - the 32-channel
  // K-block loaders hold a 3x3 halo at most
  const int cact = cin_act ? cin_act : p.cin;
  const bool c16 = tc && f16_ok && ((cact <= 16 && p.cin2 <= 16) || (p.k == 7 && cact <= 32 && p.cin2 == 0));
  pc.t16c_off = c16 ? A.put(pack_tc16(p, 0, 16)) : (size_t)-1;
  pc.cin = p.cin; pc.cin2 = p.cin2; pc.coutp = p.coutp; pc.k = p.k;
  return pc;
}

// the 19 trunk convs in execution order (see fvp_run_trunk2d / c2c_forward)
void pack_trunk(const fvp_ctx* ctx, const std::string& p, Arena& A, std::vector<PendingConv>& out, bool tc) {
  const std::string ed = p + ".encoder_decoder";
  const int JP = ctx->geom.proj.JP;
  out.push_back(stash(A, pack_conv(ctx, p + ".front_layers.0.block.0"), tc, JP));
  out.push_back(stash(A, pack_conv(ctx, p + ".front_layers.1.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, p + ".front_layers.1.res_branch.3", p + ".front_layers.1.skip_con.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".skip_res1.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".skip_res1.res_branch.3"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".encoder_res1.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".encoder_res1.res_branch.3", ed + ".encoder_res1.skip_con.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".skip_res2.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".skip_res2.res_branch.3"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".encoder_res2.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".encoder_res2.res_branch.3", ed + ".encoder_res2.skip_con.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".mid_res.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".mid_res.res_branch.3"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".decoder_res2.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".decoder_res2.res_branch.3"), tc));
  out.push_back(stash(A, pack_convT(ctx, ed + ".decoder_upsample2.block.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".decoder_res1.res_branch.0"), tc));
  out.push_back(stash(A, pack_conv(ctx, ed + ".decoder_res1.res_branch.3"), tc));
  out.push_back(stash(A, pack_convT(ctx, ed + ".decoder_upsample1.block.0"), tc));
}

FvpConvW bind(const float* base, const PendingConv& pc) {
  FvpConvW w;
  w.w = base + pc.w_off;
  w.b = base + pc.b_off;
  w.cin = pc.cin; w.cin2 = pc.cin2; w.coutp = pc.coutp; w.k = pc.k;
  for (int v = 0; v < 3; ++v) {
    w.wtc[v] = pc.tc_off[v] == (size_t)-1 ? nullptr : base + pc.tc_off[v];
    w.wtc16[v] = pc.t16_off[v] == (size_t)-1 ? nullptr : base + pc.t16_off[v];
  }
  w.wtc16_c16 = pc.t16c_off == (size_t)-1 ? nullptr : base + pc.t16c_off;
  return w;
}

void bind_trunk(const float* base, const std::vector<PendingConv>& v, FvpTrunkW& t) {
  FvpConvW* slots[19] = {&t.front, &t.r1a, &t.r1b, &t.s1a, &t.s1b, &t.e1a, &t.e1b, &t.s2a, &t.s2b, &t.e2a,
                         &t.e2b,   &t.ma,  &t.mb,  &t.d2a, &t.d2b, &t.up2, &t.d1a, &t.d1b, &t.up1};
  for (int i = 0; i < 19; ++i) *slots[i] = bind(base, v[i]);
}

}  // namespace

// Host-only test hook: the fp16 hi/lo weight image pack_tc16 builds for the tensor-core convolution, from a BN-folded GEMM
// matrix w_rows [(k*k*round_up(cin,16) + round_up(cin2,16))][coutp] (rows = (tap, input channel), then the fused skip conv's
// channels).  variant 0/1/2 = N tiles of up to 128/32/64 columns, cb = channels per K-block (32, or 16 for <= 16-channel
// layers).  Writes the image as halves; returns 0, -1 on bad arguments, -2 when `capacity` halves are not enough.
extern "C" int fvp_debug_pack_tc16(const float* w_rows, int cin, int cin2, int coutp, int k, int variant, int cb,
                                   unsigned short* out, long long capacity, long long* n_halves) {
  if (!w_rows || !n_halves || cin <= 0 || cin2 < 0 || coutp <= 0 || (k != 1 && k != 3 && k != 7) || variant < 0 || variant > 2 ||
      (cb != 16 && cb != 32))
    return -1;
  Packed p;
  p.cin = cin; p.cin2 = cin2; p.coutp = coutp; p.k = k;
  const size_t rows = (size_t)k * k * fvp_round_up(cin, 16) + (cin2 ? fvp_round_up(cin2, 16) : 0);
  p.w.assign(w_rows, w_rows + rows * coutp);
  p.b.assign(coutp, 0.f);
  const std::vector<float> img = pack_tc16(p, variant, cb);
  *n_halves = (long long)img.size() * 2;
  if (!out || capacity < *n_halves) return -2;
  std::memcpy(out, img.data(), img.size() * sizeof(float));
  return 0;
}

int fvp_pack_params(fvp_ctx* ctx) {
  for (const FvpParam& p : ctx->params)
    if (!p.set) return fvp_fail(ctx, FVP_E_STATE, "parameter '%s' was never set", p.name.c_str());
  Arena A;
  std::vector<PendingConv> cn, c2c, p2p;
  // ---- CenterNet -----------------------------------------------------------------------------
  pack_trunk(ctx, "pose_net.center_net", A, cn, true);
  {
    // both 3x3 heads as one 32->64 conv (+ReLU), both 1x1 heads as one block-diagonal 64->4 conv
    const Packed a = pack_conv(ctx, "pose_net.center_net.output_hm.0");
    const Packed b = pack_conv(ctx, "pose_net.center_net.output_size.0");
    Packed m;
    m.cin = 32; m.cin2 = 0; m.coutp = 64; m.k = 3;
    m.w.assign((size_t)9 * 32 * 64, 0.f);
    m.b.assign(64, 0.f);
    for (int r = 0; r < 9 * 32; ++r)
      for (int c = 0; c < 32; ++c) {
        m.w[(size_t)r * 64 + c] = a.w[(size_t)r * 32 + c];
        m.w[(size_t)r * 64 + 32 + c] = b.w[(size_t)r * 32 + c];
      }
    for (int c = 0; c < 32; ++c) { m.b[c] = a.b[c]; m.b[32 + c] = b.b[c]; }
    cn.push_back(stash(A, m, true));
    const std::vector<float>&wh = P(ctx, "pose_net.center_net.output_hm.2.weight"),
                      &bh = P(ctx, "pose_net.center_net.output_hm.2.bias"),
                      &ws = P(ctx, "pose_net.center_net.output_size.2.weight"),
                      &bs = P(ctx, "pose_net.center_net.output_size.2.bias");
    Packed h;
    h.cin = 64; h.cin2 = 0; h.coutp = 4; h.k = 1;
    h.w.assign((size_t)64 * 4, 0.f);
    h.b.assign(4, 0.f);
    for (int ci = 0; ci < 32; ++ci) {
      h.w[(size_t)ci * 4 + 0] = wh[ci];
      h.w[(size_t)(32 + ci) * 4 + 1] = ws[ci];
      h.w[(size_t)(32 + ci) * 4 + 2] = ws[32 + ci];
    }
    h.b[0] = bh[0]; h.b[1] = bs[0]; h.b[2] = bs[1];
    cn.push_back(stash(A, h, true));
  }
  // ---- C2CNet --------------------------------------------------------------------------------
  pack_trunk(ctx, "pose_net.c2c_net", A, c2c, false);
  c2c.push_back(stash(A, pack_conv(ctx, "pose_net.c2c_net.output_hm")));
  // ci-major copies for the split-K 1-D kernel: rows [ci][tap] (real input channels only), then the skip rows [ci]
  std::vector<size_t> c2c_cimajor;
  for (const PendingConv& pc : c2c) {
    const int cinP = fvp_round_up(pc.cin, 16), taps = pc.k;      // 1-D: taps = k
    const float* src = A.host.data() + pc.w_off;
    std::vector<float> v((size_t)(pc.cin * taps + pc.cin2) * pc.coutp, 0.f);
    for (int ci = 0; ci < pc.cin; ++ci)
      for (int tp = 0; tp < taps; ++tp)
        std::memcpy(&v[(size_t)(ci * taps + tp) * pc.coutp], src + (size_t)(tp * cinP + ci) * pc.coutp, pc.coutp * sizeof(float));
    for (int ci = 0; ci < pc.cin2; ++ci)
      std::memcpy(&v[(size_t)(pc.cin * taps + ci) * pc.coutp], src + (size_t)(taps * cinP + ci) * pc.coutp, pc.coutp * sizeof(float));
    c2c_cimajor.push_back(A.put(v));
  }
  // per-rank slices for the 8-CTA cluster form of the column kernel (FVP_C2C_CLUSTER=0 keeps one CTA per column, for A/B)
  size_t c2c_cluster_off = (size_t)-1;
  {
    const char* sw = std::getenv("FVP_C2C_CLUSTER");
    const size_t nfl = fvp_c2c_cluster_floats(ctx->cfg.num_joints);
    if (nfl > 0 && !(sw && sw[0] == '0')) {
      const float *w2h[20], *bh[20];
      for (int i = 0; i < 20; ++i) {
        w2h[i] = A.host.data() + c2c_cimajor[i];
        bh[i] = A.host.data() + c2c[i].b_off;
      }
      std::vector<float> v(nfl, 0.f);
      fvp_c2c_pack_cluster(w2h, bh, ctx->cfg.num_joints, v.data());
      c2c_cluster_off = A.put(v);
    }
  }
  // ---- P2PNet --------------------------------------------------------------------------------
  pack_trunk(ctx, "joint_net.conv_net", A, p2p, true);
  p2p.push_back(stash(A, pack_conv(ctx, "joint_net.conv_net.output_layer"), true));
  // ---- WeightNet -----------------------------------------------------------------------------
  const std::string wn = "joint_net.weight_net";
  const FvpLayer& wl = *find_layer(ctx, wn + ".heatmap_feature_net.0");
  std::vector<double> s, t;
  bn_affine(ctx, wl, s, t);
  const int F = ctx->cfg.feat_channels;
  std::vector<float> cw((size_t)F * 9), cb(F);
  {
    const std::vector<float>& w = P(ctx, wn + ".heatmap_feature_net.0.weight");
    for (int c = 0; c < F; ++c) {
      for (int i = 0; i < 9; ++i) cw[(size_t)c * 9 + i] = (float)((double)w[(size_t)c * 9 + i] * s[c]);
      cb[c] = (float)t[c];
    }
  }
  const size_t o_cw = A.put(cw), o_cb = A.put(cb);
  const size_t o_f1w = A.put(P(ctx, wn + ".output.0.weight")), o_f1b = A.put(P(ctx, wn + ".output.0.bias"));
  const size_t o_f2w = A.put(P(ctx, wn + ".output.2.weight")), o_f2b = A.put(P(ctx, wn + ".output.2.bias"));

  // ---- upload --------------------------------------------------------------------------------
  if (ctx->d_weights) cudaFree(ctx->d_weights);
  ctx->d_weights = nullptr;
  FVP_CUDA_OK(cudaMalloc(&ctx->d_weights, A.host.size() * sizeof(float)));
  FVP_CUDA_OK(cudaMemcpy(ctx->d_weights, A.host.data(), A.host.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* base = ctx->d_weights;
  bind_trunk(base, cn, ctx->w_center);
  ctx->w_center.head_a = bind(base, cn[19]);
  ctx->w_center.head_b = bind(base, cn[20]);
  bind_trunk(base, p2p, ctx->w_p2p);
  ctx->w_p2p.head_a = FvpConvW{nullptr, nullptr, 0, 0, 0, 0, {nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, nullptr};
  ctx->w_p2p.head_b = bind(base, p2p[19]);
  for (int i = 0; i < 20; ++i) {
    ctx->w_c2c.w[i] = base + c2c[i].w_off;
    ctx->w_c2c.b[i] = base + c2c[i].b_off;
    ctx->w_c2c.w2[i] = base + c2c_cimajor[i];
  }
  ctx->w_c2c.w3 = c2c_cluster_off == (size_t)-1 ? nullptr : base + c2c_cluster_off;
  ctx->w_c2c.max_clusters = fvp_c2c_max_clusters(ctx->cfg.num_joints);
  {  // chunk table of k_proposals' network-wide weight ring (device addresses of the ci-major weights)
    std::vector<char> plan(fvp_c2c_plan_bytes(), 0);
    const int nchunks = fvp_c2c_build_plan(ctx->w_c2c.w2, ctx->cfg.num_joints, plan.data());
    if (nchunks <= 0) return fvp_fail(ctx, FVP_E_INVALID, "C2CNet weight-ring plan failed (%d)", nchunks);
    if (ctx->d_c2c_plan) cudaFree(ctx->d_c2c_plan);
    ctx->d_c2c_plan = nullptr;
    FVP_CUDA_OK(cudaMalloc(&ctx->d_c2c_plan, plan.size()));
    FVP_CUDA_OK(cudaMemcpy(ctx->d_c2c_plan, plan.data(), plan.size(), cudaMemcpyHostToDevice));
    ctx->w_c2c.plan = ctx->d_c2c_plan;
  }
  ctx->w_pose.conv_w = base + o_cw;
  ctx->w_pose.conv_b = base + o_cb;
  ctx->w_pose.fc1_w = base + o_f1w;
  ctx->w_pose.fc1_b = base + o_f1b;
  ctx->w_pose.fc2_w = base + o_f2w;
  ctx->w_pose.fc2_b = base + o_f2b;
  ctx->w_pose.feat = F;
  ctx->w_pose.hidden = ctx->cfg.hidden_channels;
  ctx->fp16_fallback_layers = 0;                     // tensor-core layers whose folded weights left the fp16 range
  for (const std::vector<PendingConv>* v : {&cn, &p2p})
    for (const PendingConv& pc : *v)
      if (pc.tc_off[0] != (size_t)-1 && pc.t16_off[0] == (size_t)-1) ++ctx->fp16_fallback_layers;
  ctx->params_ready = true;
  return FVP_OK;
}

// ------------------------------------------------------------------------------------------------
// test hook: one standalone convolution (raw [Cout][Cin][k][k] weights, no BN) through either conv engine
// ------------------------------------------------------------------------------------------------
int fvp_debug_conv_impl(fvp_ctx* ctx, const float* d_in, int n, int H, int W, int cin, const float* h_w, const float* h_b,
                        int cout, int k, int relu, int mode, float* d_out, int repeat, float* ms_out, cudaStream_t st) {
  Packed p;
  const int taps = k * k, cinP = fvp_round_up(cin, 16), coutP = fvp_round_up(cout, 4);
  p.cin = cin; p.cin2 = 0; p.coutp = coutP; p.k = k;
  p.w.assign((size_t)taps * cinP * coutP, 0.f);
  p.b.assign(coutP, 0.f);
  for (int co = 0; co < cout; ++co) {
    for (int ci = 0; ci < cin; ++ci)
      for (int tp = 0; tp < taps; ++tp) p.w[(size_t)(tp * cinP + ci) * coutP + co] = h_w[((size_t)co * cin + ci) * taps + tp];
    p.b[co] = h_b[co];
  }
  Arena A;
  const bool c16 = cin <= 16 || (k == 7 && cin <= 32);               // 16-channel K-blocks, as stash() decides for the trunks
  const size_t ow = A.put(p.w), ob = A.put(p.b), ot = A.put(pack_tc(p, cin, 0)), ot16 = A.put(pack_tc16(p, 0)), ot16c = c16 ? A.put(pack_tc16(p, 0, 16)) : 0;
  float* d = nullptr;
  FVP_CUDA_OK(cudaMalloc(&d, A.host.size() * sizeof(float)));
  FVP_CUDA_OK(cudaMemcpy(d, A.host.data(), A.host.size() * sizeof(float), cudaMemcpyHostToDevice));
  FvpConvArgs a;
  a.in = d_in; a.H = H; a.W = W; a.Cin = cin; a.in2 = nullptr; a.Cin2 = 0;
  a.w = d + ow; a.bias = d + ob; a.out = d_out; a.CoutP = coutP; a.CoutS = coutP; a.CoutReal = cout;
  a.res = nullptr; a.res_mode = 0; a.relu = relu; a.ksize = k; a.upsample = 0; a.nchw = 0; a.n = n; a.valid = nullptr;
  a.fmt = 0;
  // mode 4: the TMA-fed form of engine 2 - input converted to a split tensor first, split output converted back afterwards
  // (both conversions outside the timed launches), so the same fp32 NHWC interface tests the split path in isolation
  void *d_split_in = nullptr, *d_split_out = nullptr;
  const size_t n_in = (size_t)n * H * W * cin, n_out = (size_t)n * H * W * coutP;
  if ((mode & 0xff) == 4) {
    if (cin % 16 || coutP % 16) { cudaFree(d); return fvp_fail(ctx, FVP_E_INVALID, "debug conv mode 4 needs cin and cout multiples of 16"); }
    FVP_CUDA_OK(cudaMalloc(&d_split_in, n_in * 4));
    FVP_CUDA_OK(cudaMalloc(&d_split_out, n_out * 4));
    fvp_launch_split(d_in, d_split_in, n_in, st);
    a.in = (const float*)d_split_in;
    a.out = (float*)d_split_out;
    a.fmt = FVP_FMT_IN_SPLIT | FVP_FMT_OUT_SPLIT;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const bool prof = (mode & 0x100) != 0;                            // role counters: ms_out must hold 1 + 12 floats
  mode &= 0xff;
  unsigned long long* d_prof = nullptr;
  if (prof) { FVP_CUDA_OK(cudaMalloc(&d_prof, 12 * sizeof(unsigned long long))); }
  FvpLaunchEnv env{ctx->num_sms, mode, nullptr, nullptr, 0, &ctx->launch_error};
  for (int it = 0; it < 1 + (repeat > 0 ? repeat : 1); ++it) {     // first launch = warm-up
    if (it == 1) {
      cudaEventRecord(e0, st);
      if (prof) { cudaMemsetAsync(d_prof, 0, 12 * sizeof(unsigned long long), st); env.tc_prof = d_prof; }
    }
    const float* const i16c[3] = {d + ot16c, nullptr, nullptr};
    const float* const i16[3] = {d + ot16, nullptr, nullptr};
    const float* const i32[3] = {d + ot, nullptr, nullptr};
    if (mode == 4) fvp_launch_conv_tc(a, (cin <= 16 && k != 7) ? i16c : i16, (cin <= 16 && k != 7) ? 2 : 1, env, st);
    else if (mode == 2 && c16) fvp_launch_conv_tc(a, i16c, 2, env, st);
    else if (mode == 2 || mode == 3) fvp_launch_conv_tc(a, i16, 1, env, st);
    else if (mode == 1) fvp_launch_conv_tc(a, i32, 0, env, st);
    else fvp_launch_conv(a, st);
  }
  cudaEventRecord(e1, st);
  if (d_split_out) fvp_launch_unsplit(d_split_out, d_out, n_out, st);
  FVP_CUDA_OK(cudaStreamSynchronize(st));
  if (d_split_in) cudaFree(d_split_in);
  if (d_split_out) cudaFree(d_split_out);
  if (ms_out) { cudaEventElapsedTime(ms_out, e0, e1); *ms_out /= (float)(repeat > 0 ? repeat : 1); }
  if (prof && ms_out) {                                             // kilo-cycles per launch, summed over CTAs
    unsigned long long h[12];
    cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 12; ++i) ms_out[1 + i] = (float)((double)h[i] / 1000.0 / (repeat > 0 ? repeat : 1));
  }
  if (d_prof) cudaFree(d_prof);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  FVP_CUDA_OK(cudaGetLastError());
  cudaFree(d);
  if (ctx->launch_error) { ctx->launch_error = 0; return fvp_fail(ctx, FVP_E_CUDA, "a convolution launch could not be prepared (TMA descriptor)"); }
  return FVP_OK;
}
