// fvp_project.cuh - in-kernel voxel -> heat-map-pixel projection and bilinear taps.
//
// fvp_project() is the CUDA twin of oracle/fvp_oracle.py::project_chain_np: the reference's fp32
// expression chain  project_pose -> clamp -> resize affine -> /stride -> normalise -> clamp ->
// grid_sample un-normalise  (lib/utils/cameras.py:43-54, lib/models/project_whole.py:50-59,
// lib/utils/transforms.py:59-63, ATen grid_sampler_unnormalize) with every rounding pinned by
// _rn intrinsics (which nvcc never contracts) and FMAs exactly where the reference's k=3 sgemm has
// them.  On identical fp32 inputs it is bit-identical to the reference's cached sample grid.
#pragma once
#include "fvp_common.cuh"

// Correctly rounded x / d from a correctly rounded reciprocal r = RN(1/d): q = RN(x*r); rem = x - q*d (exact, FMA);
// q' = RN(q + rem*r) (Markstein).  Three FMA-pipe instructions instead of the ~10 of __fdiv_rn's slow path; valid for
// the normal-range operands of this chain (pixel / millimetre magnitudes) - tests/test_gpu_parity.py checks the whole
// chain bit-for-bit against oracle.project_chain_np on random points.
__device__ __forceinline__ float fvp_div_by(float x, float d, float r) {
  const float q = __fmul_rn(x, r);
  const float rem = __fmaf_rn(-q, d, x);
  return __fmaf_rn(rem, r, q);
}

__device__ __forceinline__ void fvp_project(const FvpCam& c, const float* __restrict__ A, const FvpProj& P,
                                            float px, float py, float pz, float& ix, float& iy) {
  const float dx = __fsub_rn(px, c.T[0]);
  const float dy = __fsub_rn(py, c.T[1]);
  const float dz = __fsub_rn(pz, c.T[2]);
  // torch.mm(R, x.T - T): acc = R0*dx ; acc = fma(R1,dy,acc) ; acc = fma(R2,dz,acc)
  const float q0 = __fmaf_rn(c.R[2], dz, __fmaf_rn(c.R[1], dy, __fmul_rn(c.R[0], dx)));
  const float q1 = __fmaf_rn(c.R[5], dz, __fmaf_rn(c.R[4], dy, __fmul_rn(c.R[3], dx)));
  const float q2 = __fmaf_rn(c.R[8], dz, __fmaf_rn(c.R[7], dy, __fmul_rn(c.R[6], dx)));
  const float den = __fadd_rn(q2, 1e-5f);                 // no cheirality test (cameras.py:44)
  float y0, y1;
  {
    const float aden = fabsf(den);
    if (aden > 1e-3f && aden < 1e9f) {                    // common case: shared correctly-rounded reciprocal
      const float rden = __frcp_rn(den);
      y0 = fvp_div_by(q0, den, rden);
      y1 = fvp_div_by(q1, den, rden);
    } else {                                              // grazing the camera plane: full IEEE division
      y0 = __fdiv_rn(q0, den);
      y1 = __fdiv_rn(q1, den);
    }
  }
  const float r2 = __fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1));
  float d = __fadd_rn(1.0f, __fmul_rn(c.k[0], r2));
  d = __fadd_rn(d, __fmul_rn(__fmul_rn(c.k[1], r2), r2));
  d = __fadd_rn(d, __fmul_rn(__fmul_rn(__fmul_rn(c.k[2], r2), r2), r2));
  float u = __fadd_rn(__fmul_rn(y0, d), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c.p[0]), y0), y1));
  u = __fadd_rn(u, __fmul_rn(c.p[1], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, y0), y0))));
  float v = __fadd_rn(__fmul_rn(y1, d), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c.p[1]), y0), y1));
  v = __fadd_rn(v, __fmul_rn(c.p[0], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, y1), y1))));
  float X = __fadd_rn(__fmul_rn(c.fx, u), c.cx);
  float Y = __fadd_rn(__fmul_rn(c.fy, v), c.cy);
  X = fminf(fmaxf(X, -1.0f), P.ori_max);                  // same bound for x and y (project_whole.py:51)
  Y = fminf(fmaxf(Y, -1.0f), P.ori_max);
  // torch.mm(A[2x3], [X;Y;1])
  const float ax = __fadd_rn(__fmaf_rn(A[1], Y, __fmul_rn(A[0], X)), A[2]);
  const float ay = __fadd_rn(__fmaf_rn(A[4], Y, __fmul_rn(A[3], X)), A[5]);
  const float sx = fvp_div_by(__fmul_rn(ax, P.hm_w), P.img_w, P.r_img_w);
  const float sy = fvp_div_by(__fmul_rn(ay, P.hm_h), P.img_h, P.r_img_h);
  float gx = __fsub_rn(__fmul_rn(fvp_div_by(sx, P.wm1, P.r_wm1), 2.0f), 1.0f);
  float gy = __fsub_rn(__fmul_rn(fvp_div_by(sy, P.hm1, P.r_hm1), 2.0f), 1.0f);
  gx = fminf(fmaxf(gx, -1.1f), 1.1f);
  gy = fminf(fmaxf(gy, -1.1f), 1.1f);
  ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), P.wm1);   // ((g+1)/2)*(size-1), /2 is exact
  iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), P.hm1);
}

// Four bilinear taps of one sample in the zero-bordered channel-last heat map.
// off = float4 index of the NW tap's channel group 0; weights as ATen's CPU grid_sampler computes
// them (w = x - floor(x), e = 1 - w; nw = s*e ...).  Because |g| <= 1.1 the taps always fall inside
// the border (PAD >= 0.05*(size-1)+2), so out-of-image taps read zeros: no bounds tests.
struct FvpTaps {
  int off;
  float w00, w01, w10, w11;
};

__device__ __forceinline__ FvpTaps fvp_taps(const FvpProj& P, float ix, float iy) {
  FvpTaps t;
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = __fsub_rn(ix, x0f), fy = __fsub_rn(iy, y0f);
  const float ex = __fsub_rn(1.0f, fx), ey = __fsub_rn(1.0f, fy);
  // No index clamp: fvp_project's last clamp uses fminf/fmaxf, which return the non-NaN operand, so g is always
  // inside [-1.1, 1.1] (NaN and Inf included) and floor(ix) inside [-(PAD-1), size-1+PAD-2].
  const int x0 = (int)x0f, y0 = (int)y0f;
  t.off = ((y0 + P.PADY) * P.WP + (x0 + P.PADX)) * (P.JP >> 2);
  t.w00 = __fmul_rn(ey, ex);
  t.w01 = __fmul_rn(ey, fx);
  t.w10 = __fmul_rn(fy, ex);
  t.w11 = __fmul_rn(fy, fx);
  return t;
}

// The same taps in exchange form: NW offset and the two fractions; fvp_tap_weights rebuilds the four weights with the
// very operations of fvp_taps (bit-identical), so a lane can hand (off, fx, fy) to its group instead of five values.
__device__ __forceinline__ void fvp_taps_frac(const FvpProj& P, float ix, float iy, int& off, float& fx, float& fy) {
  const float x0f = floorf(ix), y0f = floorf(iy);
  fx = __fsub_rn(ix, x0f);
  fy = __fsub_rn(iy, y0f);
  off = (((int)y0f + P.PADY) * P.WP + ((int)x0f + P.PADX)) * (P.JP >> 2);
}
__device__ __forceinline__ void fvp_tap_weights(float fx, float fy, float& w00, float& w01, float& w10, float& w11) {
  const float ex = __fsub_rn(1.0f, fx), ey = __fsub_rn(1.0f, fy);
  w00 = __fmul_rn(ey, ex);
  w01 = __fmul_rn(ey, fx);
  w10 = __fmul_rn(fy, ex);
  w11 = __fmul_rn(fy, fx);
}

// base = staged heat-map buffer (kernel-uniform); off = 32-bit float4 index of the NW tap including the (frame, view)
// image offset and the thread's channel group.  Byte offsets stay unsigned 32-bit so that every tap address is
// uniform base + zero-extended register (fvp_create bounds the buffer below 4 GiB).
__device__ __forceinline__ float4 fvp_ldg_at(const float4* __restrict__ base, unsigned byte_off) {
  return __ldg((const float4*)((const char*)base + byte_off));
}
// PX16 > 0: compile-time pixel stride in bytes (lets the east taps share the west taps' address register).
template <int PX16 = 0>
__device__ __forceinline__ void fvp_tap_accumulate(float4& acc, const float4* __restrict__ base, int off,
                                                   int row_stride4, int px_stride4, float w00, float w01,
                                                   float w10, float w11) {
  const unsigned o0 = (unsigned)off << 4, o2 = o0 + ((unsigned)row_stride4 << 4);
  float4 a, b, c, d;
  if (PX16 > 0) {
    const float4* pn = (const float4*)((const char*)base + o0);
    const float4* ps = (const float4*)((const char*)base + o2);
    a = __ldg(pn); b = __ldg(pn + PX16 / 16); c = __ldg(ps); d = __ldg(ps + PX16 / 16);
  } else {
    const unsigned px = (unsigned)px_stride4 << 4;
    a = fvp_ldg_at(base, o0); b = fvp_ldg_at(base, o0 + px); c = fvp_ldg_at(base, o2); d = fvp_ldg_at(base, o2 + px);
  }
  acc.x = fmaf(d.x, w11, fmaf(c.x, w10, fmaf(b.x, w01, fmaf(a.x, w00, acc.x))));
  acc.y = fmaf(d.y, w11, fmaf(c.y, w10, fmaf(b.y, w01, fmaf(a.y, w00, acc.y))));
  acc.z = fmaf(d.z, w11, fmaf(c.z, w10, fmaf(b.z, w01, fmaf(a.z, w00, acc.z))));
  acc.w = fmaf(d.w, w11, fmaf(c.w, w10, fmaf(b.w, w01, fmaf(a.w, w00, acc.w))));
}

// Footprint cache of K3: consecutive depths of a column often fall into the same 2x2 pixel cell of a view, so the four
// float4 taps are kept in registers and reloaded by ALL lanes as soon as ANY lane left its cell - a vote + uniform branch
// instead of a divergent region (BSSY/BSYNC, serialised paths); lanes that would have hit re-read the same bytes, values
// are identical.  Must be called by converged warps.
struct FvpTapRegs {
  int off;
  float4 a, b, c, d;
};
template <int PX16 = 0>
__device__ __forceinline__ void fvp_tap_accumulate_vote(float4& acc, FvpTapRegs& tr, const float4* __restrict__ base, int off,
                                                        int row_stride4, int px_stride4, float w00, float w01, float w10,
                                                        float w11) {
  if (__any_sync(0xffffffffu, off != tr.off)) {
    const unsigned o0 = (unsigned)off << 4, o2 = o0 + ((unsigned)row_stride4 << 4);
    if (PX16 > 0) {
      const float4* pn = (const float4*)((const char*)base + o0);
      const float4* ps = (const float4*)((const char*)base + o2);
      tr.a = __ldg(pn); tr.b = __ldg(pn + PX16 / 16); tr.c = __ldg(ps); tr.d = __ldg(ps + PX16 / 16);
    } else {
      const unsigned px = (unsigned)px_stride4 << 4;
      tr.a = fvp_ldg_at(base, o0); tr.b = fvp_ldg_at(base, o0 + px); tr.c = fvp_ldg_at(base, o2); tr.d = fvp_ldg_at(base, o2 + px);
    }
  }
  tr.off = off;
  const float4 a = tr.a, b = tr.b, c = tr.c, d = tr.d;
  acc.x = fmaf(d.x, w11, fmaf(c.x, w10, fmaf(b.x, w01, fmaf(a.x, w00, acc.x))));
  acc.y = fmaf(d.y, w11, fmaf(c.y, w10, fmaf(b.y, w01, fmaf(a.y, w00, acc.y))));
  acc.z = fmaf(d.z, w11, fmaf(c.z, w10, fmaf(b.z, w01, fmaf(a.z, w00, acc.z))));
  acc.w = fmaf(d.w, w11, fmaf(c.w, w10, fmaf(b.w, w01, fmaf(a.w, w00, acc.w))));
}

// mean over V views followed by clamp(0,1): correctly rounded acc / V through one Newton correction
// (q = acc*r; q += fma(-q,V,acc)*r), then min/max.
__device__ __forceinline__ float fvp_mean_clamp(float acc, float fV, float rV) {
  float q = __fmul_rn(acc, rV);
  const float rem = __fmaf_rn(-q, fV, acc);
  q = __fmaf_rn(rem, rV, q);
  return fminf(fmaxf(q, 0.0f), 1.0f);
}
__device__ __forceinline__ float4 fvp_mean_clamp4(float4 a, float fV, float rV) {
  return make_float4(fvp_mean_clamp(a.x, fV, rV), fvp_mean_clamp(a.y, fV, rV), fvp_mean_clamp(a.z, fV, rV),
                     fvp_mean_clamp(a.w, fV, rV));
}
__device__ __forceinline__ float4 fvp_max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
