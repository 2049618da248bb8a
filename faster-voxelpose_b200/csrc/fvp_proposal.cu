// fvp_proposal.cu - everything between the CenterNet output and the JLN crops:
//   * nms2D + top-k                       (lib/core/proposal.py:13-33)
//   * z-column re-sampling (K2)           (replaces the torch.gather on the never-materialised volume,
//                                          lib/models/human_detection_net.py:92-93)
//   * C2CNet                              (lib/models/cnns_1d.py:10-132): k_proposals = whole net in one CTA per column
//                                          (throughput form; Z = 20 / 40 through the weight ring, other Z generic),
//                                          k_proposals_cluster = one 8-CTA cluster per column (latency form, Z = 20)
//   * z arg-max, ProposalLayer assembly   (human_detection_net.py:44-65,95-102)
//   * JLN crop parameters                 (lib/models/project_individual.py:110-121)
#include "fvp_kernels.h"
#include "fvp_project.cuh"

// ------------------------------------------------------------------------------------------------
// NMS + top-k: one CTA per frame.  nms = (hm == maxpool3x3(hm)) ? hm : 0 ; P rounds of block arg-max.
// Ties (only possible among exact duplicates, e.g. the zeros of suppressed cells) resolve to the
// lowest flat index - torch.topk leaves that order unspecified.
// ------------------------------------------------------------------------------------------------
struct ValIdx {
  float v;
  int i;
};
__device__ __forceinline__ ValIdx vi_best(ValIdx a, ValIdx b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

// Round 2: the P selection rounds no longer rescan the map with the whole block.  The cells are split into 32 groups (the
// warps' strided cells); each group's best sits in shared memory; a round is one warp reduction over the 32 group bests
// and one re-scan of the winner's group, all in warp 0 without block barriers.  Same order as before: value descending,
// flat index ascending among equals (the tie rule of the tests), so the result is bit-identical.
__global__ void __launch_bounds__(1024) k_nms_topk(const float* __restrict__ hm, size_t img_stride, int X, int Y,
                                                   int P, float* __restrict__ conf, int* __restrict__ flat) {
  extern __shared__ float s_nms[];               // [X*Y] map, then [ceil(X*Y/32)] keep bits
  __shared__ ValIdx s_warp[32];                  // best of every group of cells (group w = the cells warp w strides over)
  const int b = blockIdx.x, n = X * Y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const float* h = hm + (size_t)b * img_stride;
  unsigned* s_keep = reinterpret_cast<unsigned*>(s_nms + n);
  // The map is staged with independent coalesced loads first; the 3x3 neighbourhoods are then read from shared memory.
  // (Measured: staging + the one-warp selection below took the stage from 21 to 18 us; the rest is launch latency and the
  // serial chain of one CTA - ncu 15 us for 49 K warp instructions.)
  for (int i = tid; i < n; i += nthreads) s_nms[i] = h[i];
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += nthreads) {     // warp-uniform trip count: every lane takes part in the ballot
    const int i = i0 + tid;
    bool keep = false;
    if (i < n) {
      const int x = i / Y, y = i - x * Y;
      const float v = s_nms[i];
      float m = v;
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
          const int xx = x + dx, yy = y + dy;
          if (xx >= 0 && xx < X && yy >= 0 && yy < Y) m = fmaxf(m, s_nms[xx * Y + yy]);
        }
      keep = v == m;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    if (lane == 0 && i < n) s_keep[i >> 5] = bits;   // i is a multiple of 32 for lane 0
  }
  __syncthreads();
  for (int i = tid; i < n; i += nthreads)
    if (!((s_keep[i >> 5] >> (i & 31)) & 1u)) s_nms[i] = 0.0f;
  __syncthreads();
  const ValIdx none = {-INFINITY, 0x7fffffff};
  auto warp_best = [&](ValIdx v) {               // all lanes get the warp's best
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ValIdx other;
      other.v = __shfl_xor_sync(0xffffffffu, v.v, o);
      other.i = __shfl_xor_sync(0xffffffffu, v.i, o);
      v = vi_best(v, other);
    }
    return v;
  };
  auto best_of_warp_cells = [&](int w) {         // best of the cells i with (i % nthreads) / 32 == w, computed by one whole warp
    ValIdx best = none;
    for (int i = w * 32 + lane; i < n; i += nthreads) best = vi_best(best, ValIdx{s_nms[i], i});
    return warp_best(best);
  };
  {
    const ValIdx wb = best_of_warp_cells(warp);
    if (lane == 0) s_warp[warp] = wb;
  }
  __syncthreads();
  if (warp != 0) return;
  // The P rounds run in warp 0 alone, without block barriers: reduce the 32 group bests, emit the winner, remove it, re-scan
  // the winner's group (<= ceil(n / 32) cells over 32 lanes).  Same order as a global arg-max per round.
  const int nwarps = nthreads >> 5;
  for (int r = 0; r < P; ++r) {
    const ValIdx v2 = warp_best(lane < nwarps ? s_warp[lane] : none);
    if (lane == 0) {
      conf[b * P + r] = v2.v;
      flat[b * P + r] = v2.i;
      s_nms[v2.i] = -INFINITY;                   // remove from later rounds
    }
    __syncwarp();
    const int w = (v2.i % nthreads) >> 5;        // the group that owned the winner has a new best
    const ValIdx nb = best_of_warp_cells(w);
    if (lane == 0) s_warp[w] = nb;
    __syncwarp();
  }
}

static inline size_t fvp_nms_smem_bytes(int X, int Y) {           // the map + one keep bit per cell
  return (size_t)X * Y * sizeof(float) + (size_t)((X * Y + 31) / 32) * sizeof(unsigned);
}

void fvp_launch_nms_topk(const float* d_hm, size_t img_stride, int X, int Y, int P, int batch, float* d_conf,
                         int* d_flat, cudaStream_t st) {
  const size_t smem = fvp_nms_smem_bytes(X, Y);                   // opt-in size set per device by fvp_proposal_init_device
  k_nms_topk<<<batch, 1024, smem, st>>>(d_hm, img_stride, X, Y, P, d_conf, d_flat);
}

// ------------------------------------------------------------------------------------------------
// 1-D trunk in shared memory.  Activations [C][L] fp32; weights packed [(tap*CinP + ci) (+ CinP2 rows)][Cout].
// ------------------------------------------------------------------------------------------------
constexpr int C2C_THREADS = 512;
constexpr int C2C_MAXZ = 40;
constexpr int C2C_BUF = 32 * C2C_MAXZ;   // floats per activation buffer: 32xZ = 64xZ/2 = 128xZ/4

struct C2CLayer {
  const float* w;
  const float* b;
};
struct C2CNet {
  C2CLayer l[20];
};

// out[co][l] = epilogue( sum_{tap,ci} w[tap*CinP+ci][co] * in[ci][l+tap-pad]  (+ sum_ci w2[ci][co]*in2[ci][l]) + b[co] )
// upsample: GEMM columns co' = d*Co + co  ->  out[co][2l+d]
// The layer's packed weights are streamed global(L2) -> shared in 32 KB chunks with cp.async, double buffered, so the
// inner loop reads only shared memory (round 1 read every weight straight from L2 inside the FMA loop: 0.45 ms).
constexpr int C2C_WCHUNK = 8192;    // floats per weight buffer (32 KB)
constexpr int C2C_NST_MAX = 4;      // weight ring stages: 4 for Z <= 20, 3 for Z = 40 (the split-K scratch grows with Z)
static inline int c2c_nst(int Z) { return Z <= 20 ? 4 : 3; }
static inline size_t c2c_smem_bytes(int Z) {
  return (size_t)(c2c_nst(Z) * C2C_WCHUNK + 6 * C2C_BUF + 24 * C2C_MAXZ + 4 * C2C_MAXZ + C2C_THREADS * (Z <= 20 ? 20 : C2C_MAXZ)) * sizeof(float);
}

__device__ __forceinline__ void c2c_cp16(float* dst_smem, const float* src_gmem) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src_gmem));
}

__device__ void c2c_conv(const float* __restrict__ in, int Cin, int L, const float* __restrict__ in2, int Cin2,
                         C2CLayer ly, int K, float* __restrict__ out, int CoutG, const float* __restrict__ res,
                         int res_mode, bool relu, bool upsample, float* __restrict__ wbuf) {
  const int tid = threadIdx.x;
  const int CinP = (Cin + 15) & ~15;                 // 16, 32, 64 or 128: a power of two
  const int cin_shift = 31 - __clz(CinP);
  const int Cin2P = in2 ? ((Cin2 + 15) & ~15) : 0;
  const int main_rows = K * CinP, rows = main_rows + Cin2P;
  const int rpc = C2C_WCHUNK / CoutG;                 // rows per chunk
  const int nchunks = (rows + rpc - 1) / rpc;
  const int groups = C2C_THREADS / CoutG;
  const int co = tid % CoutG, lg = tid / CoutG;
  const int pad = (K - 1) / 2;
  const bool worker = lg < groups;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int lpos[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) lpos[k] = lg + k * groups;

  auto issue = [&](int ch) {
    const int r0 = ch * rpc, nr = min(rpc, rows - r0);
    const float* src = ly.w + (size_t)r0 * CoutG;
    float* dst = wbuf + (ch & 1) * C2C_WCHUNK;
    for (int i = tid * 4; i < nr * CoutG; i += C2C_THREADS * 4) c2c_cp16(dst + i, src + i);
    asm volatile("cp.async.commit_group;\n" ::);
  };
  issue(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      issue(ch + 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    if (worker) {
      const float* wb = wbuf + (ch & 1) * C2C_WCHUNK + co;
      const int r0 = ch * rpc, nr = min(rpc, rows - r0);
#pragma unroll 4
      for (int rr = 0; rr < nr; ++rr) {
        const int r = r0 + rr;
        const float* src;
        int ci, shift;
        if (r < main_rows) {
          const int tap = r >> cin_shift;
          ci = r - (tap << cin_shift);
          if (ci >= Cin) continue;                  // zero padding rows
          shift = tap - pad;
          src = in;
        } else {
          ci = r - main_rows;
          if (ci >= Cin2) continue;
          shift = 0;
          src = in2;
        }
        const float wv = wb[rr * CoutG];
        const float* xr = src + ci * L;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int p = lpos[k] + shift;
          const float x = (lpos[k] < L && p >= 0 && p < L) ? xr[p] : 0.f;
          acc[k] = fmaf(wv, x, acc[k]);
        }
      }
    }
    __syncthreads();                                   // buffer (ch&1) may be refilled by chunk ch+2
  }
  if (worker) {
    const float bias = __ldg(ly.b + co);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int l = lpos[k];
      if (l >= L) continue;
      float v = acc[k] + bias;
      int oc = co, ol = l, Lo = L;
      if (upsample) {
        const int Co = CoutG >> 1;
        const int d = co / Co;
        oc = co - d * Co;
        ol = 2 * l + d;
        Lo = 2 * L;
      }
      if (res_mode == 1) v += res[oc * Lo + ol];
      if (relu) v = fmaxf(v, 0.f);
      if (res_mode == 2) v += res[oc * Lo + ol];
      out[oc * Lo + ol] = v;
    }
  }
  __syncthreads();
}

__device__ void c2c_pool(const float* __restrict__ in, int C, int L, float* __restrict__ out) {
  const int Lo = L >> 1;
  for (int i = threadIdx.x; i < C * Lo; i += C2C_THREADS) {
    const int c = i / Lo, l = i - c * Lo;
    out[i] = fmaxf(in[c * L + 2 * l], in[c * L + 2 * l + 1]);
  }
  __syncthreads();
}

// x0: [>=J][Z] input columns; result hm1d[Z] lands in `out` (smem)
__device__ void c2c_forward(const C2CNet& net, const float* x0, int J, int Z, float* buf, float* out, float* wbuf) {
  float *B0 = buf, *B1 = buf + C2C_BUF, *B2 = buf + 2 * C2C_BUF, *B3 = buf + 3 * C2C_BUF, *B4 = buf + 4 * C2C_BUF,
        *B5 = buf + 5 * C2C_BUF;
  const int L = Z, L2 = Z / 2, L4 = Z / 4;
  c2c_conv(x0, J, L, nullptr, 0, net.l[0], 7, B0, 16, nullptr, 0, true, false, wbuf);
  c2c_conv(B0, 16, L, nullptr, 0, net.l[1], 3, B1, 32, nullptr, 0, true, false, wbuf);
  c2c_conv(B1, 32, L, B0, 16, net.l[2], 3, B2, 32, nullptr, 0, true, false, wbuf);
  c2c_conv(B2, 32, L, nullptr, 0, net.l[3], 3, B0, 32, nullptr, 0, true, false, wbuf);
  c2c_conv(B0, 32, L, nullptr, 0, net.l[4], 3, B3, 32, B2, 1, true, false, wbuf);          // skip1
  c2c_pool(B2, 32, L, B0);
  c2c_conv(B0, 32, L2, nullptr, 0, net.l[5], 3, B1, 64, nullptr, 0, true, false, wbuf);
  c2c_conv(B1, 64, L2, B0, 32, net.l[6], 3, B4, 64, nullptr, 0, true, false, wbuf);         // e1
  c2c_conv(B4, 64, L2, nullptr, 0, net.l[7], 3, B0, 64, nullptr, 0, true, false, wbuf);
  c2c_conv(B0, 64, L2, nullptr, 0, net.l[8], 3, B5, 64, B4, 1, true, false, wbuf);          // skip2
  c2c_pool(B4, 64, L2, B0);
  c2c_conv(B0, 64, L4, nullptr, 0, net.l[9], 3, B1, 128, nullptr, 0, true, false, wbuf);
  c2c_conv(B1, 128, L4, B0, 64, net.l[10], 3, B2, 128, nullptr, 0, true, false, wbuf);      // e2
  c2c_conv(B2, 128, L4, nullptr, 0, net.l[11], 3, B0, 128, nullptr, 0, true, false, wbuf);
  c2c_conv(B0, 128, L4, nullptr, 0, net.l[12], 3, B1, 128, B2, 1, true, false, wbuf);       // m
  c2c_conv(B1, 128, L4, nullptr, 0, net.l[13], 3, B0, 128, nullptr, 0, true, false, wbuf);
  c2c_conv(B0, 128, L4, nullptr, 0, net.l[14], 3, B2, 128, B1, 1, true, false, wbuf);       // d2
  c2c_conv(B2, 128, L4, nullptr, 0, net.l[15], 1, B0, 128, B5, 2, true, true, wbuf);        // u2 = relu(convT)+skip2
  c2c_conv(B0, 64, L2, nullptr, 0, net.l[16], 3, B1, 64, nullptr, 0, true, false, wbuf);
  c2c_conv(B1, 64, L2, nullptr, 0, net.l[17], 3, B2, 64, B0, 1, true, false, wbuf);         // d1
  c2c_conv(B2, 64, L2, nullptr, 0, net.l[18], 1, B0, 64, B3, 2, true, true, wbuf);          // u1 = relu(convT)+skip1
  c2c_conv(B0, 32, L, nullptr, 0, net.l[19], 1, out, 4, nullptr, 0, false, false, wbuf);    // head (row 0 of 4)
}

// ---- network-wide weight ring -------------------------------------------------------------------------------------
// The ~1.5 MB of C2CNet weights stream through every column's CTA once.  Round 1 restarted a cp.async double buffer
// in every (layer, part): 26 exposed L2 latencies plus 2 block barriers per 32 KB chunk - 94 us per column for 2.5 MMAC.
// Now the whole net is ONE sequence of chunks (host-built table, consumption order) that thread 0 keeps NST - 1 chunks
// ahead of the consumers with TMA bulk copies (cp.async.bulk -> mbarrier complete_tx), across layer boundaries: a
// layer's first chunk is already in shared memory when the previous layer's epilogue ends.
struct C2CChunk {
  const float* src;
  int nfloats;
  int pad;
};
constexpr int C2C_MAX_CHUNKS = 80;
struct C2CPlan {
  C2CChunk chunk[C2C_MAX_CHUNKS];
  int n;
};
struct C2CRing {
  float* stage;           // [nst][C2C_WCHUNK]
  uint64_t* full;         // [nst] mbarriers
  int nst;
  int next;               // chunk the consumers wait for next (uniform across the block)
};
__device__ __forceinline__ unsigned c2c_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c2c_ring_issue(const C2CPlan& plan, const C2CRing& r, int g) {     // one thread
  if (g >= plan.n) return;
  const int st = g % r.nst;
  const unsigned bytes = (unsigned)plan.chunk[g].nfloats * 4u, bar = c2c_s32(r.full + st);
  asm volatile("{\n\t.reg .b64 s;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 s, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   c2c_s32(r.stage + (size_t)st * C2C_WCHUNK)),
               "l"(plan.chunk[g].src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ const float* c2c_ring_wait(const C2CRing& r) {                          // all threads
  const int st = r.next % r.nst;
  const unsigned bar = c2c_s32(r.full + st), parity = (unsigned)(r.next / r.nst) & 1u;
  unsigned ok, spins = 0;
  do {
    if (++spins > (1u << 26)) {          // watchdog: a protocol bug must abort the kernel, never hang the GPU
      printf("k_proposals: weight ring wait timed out (block %d thread %d chunk %d)\n", blockIdx.x, threadIdx.x, r.next);
      __trap();
    }
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
  return r.stage + (size_t)st * C2C_WCHUNK;
}
// after the block barrier that ends the use of chunk r.next: refill its stage with chunk r.next + nst
__device__ __forceinline__ void c2c_ring_release(const C2CPlan& plan, C2CRing& r) {
  if (threadIdx.x == 0) c2c_ring_issue(plan, r, r.next + r.nst);
  ++r.next;
}

// ------------------------------------------------------------------------------------------------
// Register-tiled 1-D conv with K split across the thread block (round-1 profile: the row-major c2c_conv above
// spends ~20 instructions per weight row and predicated position; 0.35 ms per column).  Thread (co, ks) owns ALL L
// output positions of channel co for the input channels ci = ks, ks+KS, ...: per ci it loads the L+K-1 inputs once
// into registers (warp-broadcast) and runs K*L FMAs; the KS partial sums meet in shared memory in fixed order.
// Weights are packed ci-major ([ci][tap][CoutG], then the fused 1x1 skip rows [ci][CoutG]) and arrive through the
// network-wide ring above; the chunking (input channels per chunk) here must match c2c_build_plan on the host.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int c2c_cpc(int C, int taps, int CoutG) {      // input channels per weight chunk
  int cpc = C2C_WCHUNK / (taps * CoutG);
  return cpc > C ? C : cpc;
}

template <int L, int K>
__device__ void c2c_conv2(const float* __restrict__ in, int Cin, const float* __restrict__ in2, int Cin2,
                          const float* bias_p, float* __restrict__ out, int CoutG, const float* __restrict__ res,
                          int res_mode, bool relu, bool upsample, const C2CPlan& plan, C2CRing& ring, float* __restrict__ red) {
  constexpr int PAD = (K - 1) / 2;
  const int tid = threadIdx.x;
  const int KS = C2C_THREADS / CoutG;
  const int co = tid % CoutG, ks = tid / CoutG;
  float acc[L];
#pragma unroll
  for (int l = 0; l < L; ++l) acc[l] = 0.f;

  for (int part = 0; part < 2; ++part) {                 // 0: main K-tap conv on `in`, 1: fused 1x1 on `in2`
    const float* src = part == 0 ? in : in2;
    if (src == nullptr) break;
    const int C = part == 0 ? Cin : Cin2;
    const int taps = part == 0 ? K : 1;
    const int cpc = c2c_cpc(C, taps, CoutG);
    const int nchunks = (C + cpc - 1) / cpc;
    for (int ch = 0; ch < nchunks; ++ch) {
      const float* wb = c2c_ring_wait(ring) + co;
      const int c_lo = ch * cpc, c_hi = min(C, c_lo + cpc);
      for (int ci = c_lo + ks; ci < c_hi; ci += KS) {
        const float* xr_src = src + ci * L;
        const float* wr = wb + (size_t)(ci - c_lo) * taps * CoutG;
        if (part == 0) {
          float xr[L + 2 * PAD];
#pragma unroll
          for (int i = 0; i < L + 2 * PAD; ++i) xr[i] = (i >= PAD && i < L + PAD) ? xr_src[i - PAD] : 0.f;
#pragma unroll
          for (int tp = 0; tp < K; ++tp) {
            const float wv = wr[tp * CoutG];
#pragma unroll
            for (int l = 0; l < L; ++l) acc[l] = fmaf(wv, xr[l + tp], acc[l]);
          }
        } else {
          const float wv = wr[0];
#pragma unroll
          for (int l = 0; l < L; ++l) acc[l] = fmaf(wv, xr_src[l], acc[l]);
        }
      }
      __syncthreads();                                   // every thread is done with this chunk's stage
      c2c_ring_release(plan, ring);
    }
  }
  // ---- fixed-order reduction over the K split, then the epilogue ----
#pragma unroll
  for (int l = 0; l < L; ++l) red[(ks * CoutG + co) * L + l] = acc[l];
  __syncthreads();
  for (int e = tid; e < CoutG * L; e += C2C_THREADS) {
    const int c = e / L, l = e - c * L;
    float v = 0.f;
    for (int k2 = 0; k2 < KS; ++k2) v += red[(k2 * CoutG + c) * L + l];
    v += __ldg(bias_p + c);
    int oc = c, ol = l, Lo = L;
    if (upsample) {
      const int Co = CoutG >> 1;
      const int d = c / Co;
      oc = c - d * Co;
      ol = 2 * l + d;
      Lo = 2 * L;
    }
    if (res_mode == 1) v += res[oc * Lo + ol];
    if (relu) v = fmaxf(v, 0.f);
    if (res_mode == 2) v += res[oc * Lo + ol];
    out[oc * Lo + ol] = v;
  }
  __syncthreads();
}

template <int L>
__device__ void c2c_pool2(const float* __restrict__ in, int C, float* __restrict__ out) {
  constexpr int Lo = L / 2;
  for (int i = threadIdx.x; i < C * Lo; i += C2C_THREADS) {
    const int c = i / Lo, l = i - c * Lo;
    out[i] = fmaxf(in[c * L + 2 * l], in[c * L + 2 * l + 1]);
  }
  __syncthreads();
}

struct C2CNet2 {          // bias per layer (the ci-major weights arrive through the ring, see C2CPlan)
  const float* b[20];
};
// (layer index, K, Cin, Cin2, CoutG) of the 20 convolutions in execution order: the single source of truth for the chunk
// plan on the host and the call sequence in c2c_forward2
struct C2CLayerShape { int K, Cin, Cin2, CoutG; };
__host__ __device__ inline C2CLayerShape c2c_shape(int i, int J) {
  switch (i) {
    case 0: return {7, J, 0, 16};
    case 1: return {3, 16, 0, 32};
    case 2: return {3, 32, 16, 32};
    case 3: case 4: return {3, 32, 0, 32};
    case 5: return {3, 32, 0, 64};
    case 6: return {3, 64, 32, 64};
    case 7: case 8: return {3, 64, 0, 64};
    case 9: return {3, 64, 0, 128};
    case 10: return {3, 128, 64, 128};
    case 11: case 12: case 13: case 14: return {3, 128, 0, 128};
    case 15: return {1, 128, 0, 128};
    case 16: case 17: return {3, 64, 0, 64};
    case 18: return {1, 64, 0, 64};
    default: return {1, 32, 0, 4};
  }
}

template <int Z>
__device__ void c2c_forward2(const C2CNet2& n, const C2CPlan& plan, C2CRing& ring, const float* x0, int J, float* buf, float* out,
                             float* red) {
  float *B0 = buf, *B1 = buf + C2C_BUF, *B2 = buf + 2 * C2C_BUF, *B3 = buf + 3 * C2C_BUF, *B4 = buf + 4 * C2C_BUF,
        *B5 = buf + 5 * C2C_BUF;
  constexpr int L = Z, L2 = Z / 2, L4 = Z / 4;
#define CV(Lx, Kx, i, inp, cin, inp2, cin2, outp, cg, resp, rm, up) \
  c2c_conv2<Lx, Kx>(inp, cin, inp2, cin2, n.b[i], outp, cg, resp, rm, true, up, plan, ring, red)
  CV(L, 7, 0, x0, J, nullptr, 0, B0, 16, nullptr, 0, false);
  CV(L, 3, 1, B0, 16, nullptr, 0, B1, 32, nullptr, 0, false);
  CV(L, 3, 2, B1, 32, B0, 16, B2, 32, nullptr, 0, false);
  CV(L, 3, 3, B2, 32, nullptr, 0, B0, 32, nullptr, 0, false);
  CV(L, 3, 4, B0, 32, nullptr, 0, B3, 32, B2, 1, false);           // skip1
  c2c_pool2<L>(B2, 32, B0);
  CV(L2, 3, 5, B0, 32, nullptr, 0, B1, 64, nullptr, 0, false);
  CV(L2, 3, 6, B1, 64, B0, 32, B4, 64, nullptr, 0, false);         // e1
  CV(L2, 3, 7, B4, 64, nullptr, 0, B0, 64, nullptr, 0, false);
  CV(L2, 3, 8, B0, 64, nullptr, 0, B5, 64, B4, 1, false);          // skip2
  c2c_pool2<L2>(B4, 64, B0);
  CV(L4, 3, 9, B0, 64, nullptr, 0, B1, 128, nullptr, 0, false);
  CV(L4, 3, 10, B1, 128, B0, 64, B2, 128, nullptr, 0, false);      // e2
  CV(L4, 3, 11, B2, 128, nullptr, 0, B0, 128, nullptr, 0, false);
  CV(L4, 3, 12, B0, 128, nullptr, 0, B1, 128, B2, 1, false);       // m
  CV(L4, 3, 13, B1, 128, nullptr, 0, B0, 128, nullptr, 0, false);
  CV(L4, 3, 14, B0, 128, nullptr, 0, B2, 128, B1, 1, false);       // d2
  CV(L4, 1, 15, B2, 128, nullptr, 0, B0, 128, B5, 2, true);        // u2 = relu(convT)+skip2
  CV(L2, 3, 16, B0, 64, nullptr, 0, B1, 64, nullptr, 0, false);
  CV(L2, 3, 17, B1, 64, nullptr, 0, B2, 64, B0, 1, false);         // d1
  CV(L2, 1, 18, B2, 64, nullptr, 0, B0, 64, B3, 2, true);          // u1 = relu(convT)+skip1
#undef CV
  c2c_conv2<L, 1>(B0, 32, nullptr, 0, n.b[19], out, 4, nullptr, 0, false, false, plan, ring, red);   // head (row 0 of 4)
}

// ------------------------------------------------------------------------------------------------
// crop parameters of one proposal (project_individual.py:110-121), exact fp32 op order
// ------------------------------------------------------------------------------------------------
__device__ void fvp_make_person(const FvpPropArgs& a, const float* c7, int seq, FvpPerson& pd) {
  pd.valid = c7[3] >= 0.0f;
  pd.seq = seq;
  pd.pad = 0;
  bool empty = false;
  for (int d = 0; d < 3; ++d) {
    const int fine = a.g.fine[d];
    const int tl = (int)rintf(__fadd_rn(__fmul_rn(c7[d], a.jln_scale[d]), a.jln_bias[d]));
    float off = __fdiv_rn((float)tl, (float)(fine - 1));
    off = __fmul_rn(off, a.whole[d]);
    off = __fsub_rn(off, __fmul_rn(a.whole[d], 0.5f));
    off = __fadd_rn(off, __fmul_rn(a.ind[d], 0.5f));
    int m = 0;
    if (d < 2) {
      m = (int)__fmul_rn(__fmul_rn(__fsub_rn(1.0f, c7[5 + d]), 0.5f), (float)(a.ind_vox[d] - 1));
      if (m < 0) m = 0;
    }
    const int start = tl + m >= 0 ? tl + m : 0;
    const int end = tl + a.ind_vox[d] - m <= fine ? tl + a.ind_vox[d] - m : fine;
    if (start >= end) empty = true;
    pd.tl[d] = tl;
    pd.offset[d] = off;
    pd.lo[d] = start - tl;
    pd.hi[d] = end - tl;
  }
  pd.empty = empty;
}

// ---- K2: the z column of the selected (x,y) cell sampled straight from the heat maps: dst[j * pitch + z], j < 4 * JG --------
__device__ void c2c_sample_column(const FvpPropArgs& a, const FvpSeq& s_seq, int flat, int b, float* __restrict__ dst, int pitch) {
  const FvpGeom& g = a.g;
  const int Z = g.Z;
  const int cx = flat / g.Y, cy = flat - cx * g.Y;       // true layout of the flattened (X,Y) map
  const float wx = g.coarse_axes[cx], wy = g.coarse_axes[g.X + cy];
  const int V = g.V;
  const float fV = (float)V, rV = 1.0f / fV;
  const int row4 = g.proj.WP * g.JG, px4 = g.JG;
  for (int i = threadIdx.x; i < Z * g.JG; i += blockDim.x) {
    const int z = i / g.JG, s = i - z * g.JG;
    const float wz = g.coarse_axes[g.X + g.Y + z];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* hm_b = (const float4*)a.hm_cl + (size_t)b * V * g.view_stride4 + s;
    for (int v = 0; v < V; ++v) {
      float ix, iy;
      fvp_project(s_seq.cam[v], s_seq.A, g.proj, wx, wy, wz, ix, iy);
      const FvpTaps t = fvp_taps(g.proj, ix, iy);
      fvp_tap_accumulate(acc, hm_b + (size_t)v * g.view_stride4, t.off, row4, px4, t.w00, t.w01, t.w10, t.w11);
    }
    const float4 val = fvp_mean_clamp4(acc, fV, rV);
    dst[(4 * s + 0) * pitch + z] = val.x;
    dst[(4 * s + 1) * pitch + z] = val.y;
    dst[(4 * s + 2) * pitch + z] = val.z;
    dst[(4 * s + 3) * pitch + z] = val.w;
  }
}

// ---- topk(1) over z, ProposalLayer (human_detection_net.py:44-65,95-102); one thread.  c2 = the cell's 2-D confidence,
// sz_w / sz_h = its bounding-box size, seq = the frame's calibration slot (the callers fetch them early) ---------------------
__device__ void c2c_emit_proposal(const FvpPropArgs& a, int slot, int seq, int flat, const float* __restrict__ s_out, float c2,
                                  float sz_w, float sz_h) {
  const FvpGeom& g = a.g;
  const int Z = g.Z;
  int iz = 0;
  float c1 = s_out[0];
  for (int z = 1; z < Z; ++z)
    if (s_out[z] > c1) { c1 = s_out[z]; iz = z; }
  const float conf = __fmul_rn(c2, c1);
  // get_index2D divides by shape[1] of the [1,X,Y] map, i.e. X (core/proposal.py:16-17,32)
  const int ixq = flat / g.X, iyq = flat - ixq * g.X;
  float c7[7];
  c7[0] = __fadd_rn(__fmul_rn((float)ixq, a.hdn_scale[0]), a.hdn_bias[0]);
  c7[1] = __fadd_rn(__fmul_rn((float)iyq, a.hdn_scale[1]), a.hdn_bias[1]);
  c7[2] = __fadd_rn(__fmul_rn((float)iz, a.hdn_scale[2]), a.hdn_bias[2]);
  c7[3] = (conf > a.min_score) ? 0.0f : -1.0f;
  c7[4] = conf;
  c7[5] = sz_w;
  c7[6] = sz_h;
  if (a.centers)
    for (int i = 0; i < 7; ++i) a.centers[(size_t)slot * 7 + i] = c7[i];
  if (a.people) {
    FvpPerson pd;
    fvp_make_person(a, c7, seq, pd);
    a.people[slot] = pd;
  }
  if (a.img_valid)
    for (int q = 0; q < 3; ++q) a.img_valid[q * a.n_slots + slot] = c7[3] >= 0.0f;
}

// one CTA per proposal slot
__global__ void __launch_bounds__(C2C_THREADS) k_proposals(FvpPropArgs a, C2CNet net, C2CNet2 net2, const C2CPlan* __restrict__ plan_g,
                                                           int nst) {
  extern __shared__ __align__(128) float c2c_smem[];
  __shared__ C2CPlan s_plan;
  __shared__ uint64_t s_full[4];
  float* s_wbuf = c2c_smem;                       // nst x C2C_WCHUNK weight ring (>= 2: the generic path double-buffers in it)
  float* s_buf = s_wbuf + nst * C2C_WCHUNK;       // 6 activation buffers
  float* s_x0 = s_buf + 6 * C2C_BUF;              // [JP][Z] input columns
  float* s_out = s_x0 + 24 * C2C_MAXZ;            // [4][Z] head output
  float* s_red = s_out + 4 * C2C_MAXZ;            // [KS][CoutG][L] split-K partial sums (512 * L floats)
  __shared__ FvpSeq s_seq;
  const FvpGeom& g = a.g;
  const int slot = blockIdx.x, b = slot / g.P;
  const int tid = threadIdx.x;
  const int J = g.J, Z = g.Z, JP = g.proj.JP;
  int flat = 0;
  // weight ring: copy the chunk table, arm the barriers, start the first nst chunks - all under the column sampling below
  C2CRing ring{s_wbuf, s_full, nst, 0};
  if (Z == 20 || Z == 40) {
    for (int i = tid; i < (int)(sizeof(C2CPlan) / 4); i += C2C_THREADS) ((int*)&s_plan)[i] = ((const int*)plan_g)[i];
    if (tid == 0) {
      for (int i = 0; i < nst; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(c2c_s32(s_full + i)));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
      for (int i = 0; i < nst; ++i) c2c_ring_issue(s_plan, ring, i);
  }

  if (a.mode == 0) {
    // ---- K2: sample the z column of the selected (x,y) cell straight from the heat maps ----------
    flat = a.flat[slot];
    const int seq = a.frame_seq[b];
    {
      const int* src = (const int*)(g.seqs + seq);
      int* dst = (int*)&s_seq;
      for (int i = tid; i < (int)(sizeof(FvpSeq) / 4); i += C2C_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    c2c_sample_column(a, s_seq, flat, b, s_x0, Z);
  } else {
    for (int i = tid; i < J * Z; i += C2C_THREADS) s_x0[i] = a.cols_in[(size_t)slot * J * Z + i];
  }
  __syncthreads();
  if (a.cols_out)
    for (int i = tid; i < J * Z; i += C2C_THREADS) a.cols_out[(size_t)slot * J * Z + i] = s_x0[i];
  (void)JP;

  if (Z == 20) c2c_forward2<20>(net2, s_plan, ring, s_x0, J, s_buf, s_out, s_red);        // s_out[0..Z) = 1-D heat map
  else if (Z == 40) c2c_forward2<40>(net2, s_plan, ring, s_x0, J, s_buf, s_out, s_red);
  else c2c_forward(net, s_x0, J, Z, s_buf, s_out, s_wbuf);

  if (a.hm1d_out)
    for (int i = tid; i < Z; i += C2C_THREADS) a.hm1d_out[(size_t)slot * Z + i] = s_out[i];
  if (a.mode != 0 || tid != 0) return;

  const float* sz = a.size + (size_t)b * a.size_img_stride + flat;
  c2c_emit_proposal(a, slot, a.frame_seq[b], flat, s_out, a.conf2d[slot], sz[0], sz[(size_t)g.X * g.Y]);
}

// ------------------------------------------------------------------------------------------------------------------------
// Cluster form (round 2): one 8-CTA thread-block cluster per column.  k_proposals above runs a column's 2.5 MMAC on ONE SM
// (92 us: shared-memory-wavefront bound, 10 of 148 SMs busy at batch 1).  Here CTA `rank` of the cluster owns 1/8 of every
// layer's GEMM columns (output channels) and keeps ITS slice of the whole network's weights and biases resident in shared
// memory (193 KB, fetched once per CTA with one TMA bulk copy per layer; a CTA that serves several columns reuses them).
// Every CTA holds a full copy of all activations.  A layer is 4 warp tasks per CTA (one per scheduler): a task is a tile of
// 4 output channels x 5 positions with the input channels across the lanes - per input channel 7 scalar + 3 vector shared
// loads feed 60 FMAs.  The 32 partial sums per output meet through a small shared scratch in fixed order, and the results
// go to all 8 CTAs' buffers with st.async (distributed shared memory): the stores complete transaction bytes on the
// DESTINATION's per-layer mbarrier, so a layer ends when a CTA has received all of its CoutG x L values - no cluster barrier
// and no fence in the layer loop (barrier.cluster per layer cost MEMBAR + ERRBAR + UCGABAR_WAIT = 40 % of the time).
//   weight slice of layer i, rank r:  [g][ci][tap][4], the fused 1x1 skip rows [g][ci2][4], the biases [g][4]
//   (g = group of 4 of the rank's CoutL = CoutG / 8 channels, zero padded; a lane reads K consecutive float4: conflict-free)
//   activations [L][C]: position-major, so the lanes (input channels) read consecutive words and a tile's 4 output
//   channels of one position are one 16-byte st.async
// Hazards: a CTA can run at most one layer ahead of the slowest one (it needs everybody's previous outputs), and a layer's
// output buffer is never an input of that layer or of the one before it, so remote stores never overwrite live data; the
// head reads B4 (written by the last trunk layer), which the next column overwrites only at its 7th layer.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int C2CL_CLUSTER = 8;
constexpr int C2CL_THREADS = 128;                  // 4 warps: one tile task per scheduler
constexpr int C2CL_TILE = 5;                       // positions per warp task; Z = 20 -> L = 20, 10, 5
constexpr int C2CL_Z = 20;
constexpr int C2CL_BUF = 640;                      // 32 x 20 = 64 x 10 = 128 x 5 floats, [L][C]
constexpr int C2CL_X0 = 24 * C2CL_Z;               // [Z][4 JG] input columns
constexpr int C2CL_SCRATCH = (C2CL_THREADS / 32) * 32 * 4 * C2CL_TILE;     // per warp: 32 lanes x 20 partial sums
constexpr int C2CL_NBAR = 20;                      // arrival barriers: 19 trunk layers + the sampled input columns
__host__ __device__ inline int c2cl_layer_floats(int i, int J) {          // 0 for the head (4 columns: rank 0 reads it from L2)
  const C2CLayerShape sh = c2c_shape(i, J);
  const int G = (sh.CoutG / C2CL_CLUSTER + 3) / 4;
  return G * 4 * (sh.K * sh.Cin + sh.Cin2 + 1);
}
__host__ __device__ inline int c2cl_slice_floats(int J) {
  int n = 0;
  for (int i = 0; i < 19; ++i) n += c2cl_layer_floats(i, J);
  return n;
}
static inline size_t c2cl_smem_bytes(int J) {
  return (size_t)(c2cl_slice_floats(J) + 6 * C2CL_BUF + C2CL_X0 + C2CL_SCRATCH + 32) * sizeof(float);
}

__device__ __forceinline__ void c2cl_wait(uint64_t* bar, unsigned parity) {
  const unsigned b = c2c_s32(bar);
  unsigned ok, spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();  // watchdog: a protocol bug must abort the kernel, never hang the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(b), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void c2cl_expect(uint64_t* bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 s;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 s, [%0], %1;\n\t}\n" ::"r"(c2c_s32(bar)), "r"(bytes) : "memory");
}
// 2 or 4 consecutive floats into the same shared-memory location of the CTAs dst0 .. dst0 + n - 1 of the cluster, counted on
// the destination's barrier (one DSMEM packet per store instead of one per value)
template <int N>
__device__ __forceinline__ void c2cl_send(const float* local, uint64_t* bar, float4 v, unsigned dst0, unsigned n) {
  const unsigned la = c2c_s32(local), lb = c2c_s32(bar);
  for (unsigned r = dst0; r < dst0 + n; ++r) {
    unsigned ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(la), "r"(r));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rb) : "r"(lb), "r"(r));
    if (N == 4)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(ra),
                   "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rb) : "memory");
    else
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];\n" ::"r"(ra), "f"(v.x),
                   "f"(v.y), "r"(rb) : "memory");
  }
}

// one layer; `abar` = this CTA's arrival barrier of the layer (expects CoutG x L x 4 bytes from the 8 ranks).  Always inlined
// with literal shapes: looped, non-inlined variants (one routine per kernel size, shapes at run time, 6 KB of code each)
// executed 58 % more instructions and were slower, at batch 1 and in the multi-column steady state.
//   in  [L][inC]  (inC >= Cin: the channel pitch), in2 [L][Cin2], out / res [Lo][Co]   - position-major, channels contiguous
template <int L, int K>
__device__ __forceinline__ void c2cl_conv(const float* __restrict__ in, int inC, int Cin, const float* __restrict__ in2, int Cin2,
                                          const float* __restrict__ w, uint64_t* wbar, uint64_t* abar, unsigned parity,
                                          float* __restrict__ out, int CoutG, const float* __restrict__ res, int res_mode,
                                          bool upsample, unsigned rank, float* __restrict__ scratch) {
  constexpr int PAD = (K - 1) / 2, TILES = L / C2CL_TILE, NX = C2CL_TILE + 2 * PAD, NV = 4 * C2CL_TILE;
  static_assert(L % C2CL_TILE == 0, "positions per layer must be a multiple of the warp tile");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CoutL = CoutG / C2CL_CLUSTER, G = (CoutL + 3) >> 2;
  const int Co = upsample ? CoutG >> 1 : CoutG;      // channels (= pitch) of the output and of the residual
  const float4* w4 = reinterpret_cast<const float4*>(w);
  c2cl_wait(wbar, 0);                                // the weights arrive once per CTA (phase 0 stays complete)
  for (int task = warp; task < G * TILES; task += C2CL_THREADS / 32) {
    const int g = task / TILES, l0 = (task - g * TILES) * C2CL_TILE;
    float acc[C2CL_TILE][4];
#pragma unroll
    for (int l = 0; l < C2CL_TILE; ++l)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
    const float4* wg = w4 + (size_t)g * Cin * K;
#pragma unroll 1
    for (int ci = lane; ci < Cin; ci += 32) {
      float xr[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const int p = l0 - PAD + i;
        xr[i] = (p >= 0 && p < L) ? in[p * inC + ci] : 0.f;
      }
#pragma unroll
      for (int tp = 0; tp < K; ++tp) {
        const float4 wv = wg[ci * K + tp];
#pragma unroll
        for (int l = 0; l < C2CL_TILE; ++l) {
          acc[l][0] = fmaf(wv.x, xr[l + tp], acc[l][0]);
          acc[l][1] = fmaf(wv.y, xr[l + tp], acc[l][1]);
          acc[l][2] = fmaf(wv.z, xr[l + tp], acc[l][2]);
          acc[l][3] = fmaf(wv.w, xr[l + tp], acc[l][3]);
        }
      }
    }
    const float4* ws = w4 + (size_t)G * Cin * K;
    if (in2 != nullptr) {
#pragma unroll 1
      for (int ci = lane; ci < Cin2; ci += 32) {
        const float4 wv = ws[g * Cin2 + ci];
#pragma unroll
        for (int l = 0; l < C2CL_TILE; ++l) {
          const float x = in2[(l0 + l) * Cin2 + ci];
          acc[l][0] = fmaf(wv.x, x, acc[l][0]);
          acc[l][1] = fmaf(wv.y, x, acc[l][1]);
          acc[l][2] = fmaf(wv.z, x, acc[l][2]);
          acc[l][3] = fmaf(wv.w, x, acc[l][3]);
        }
      }
    }
    // the 32 lanes' partial sums of the 20 outputs meet in the warp's scratch: lane j < 20 adds column j = 4 l + q in fixed order
    float* sc = scratch + warp * 32 * NV;
    {
      float4* dst = reinterpret_cast<float4*>(sc + lane * NV);
#pragma unroll
      for (int l = 0; l < C2CL_TILE; ++l) dst[l] = make_float4(acc[l][0], acc[l][1], acc[l][2], acc[l][3]);
    }
    __syncwarp();
    float s0 = 0.f, s1 = 0.f;
    if (lane < NV) {
#pragma unroll 4
      for (int r = 0; r < 32; r += 2) {
        s0 += sc[r * NV + lane];
        s1 += sc[(r + 1) * NV + lane];
      }
    }
    const float sum = s0 + s1;
    // every lane of a quad gets the 4 channels of its position; lane (l, q) sends them to the CTAs 2q and 2q + 1
    const int qb = lane & ~3, l = lane >> 2, q = lane & 3;
    float4 v;
    v.x = __shfl_sync(0xffffffffu, sum, qb);
    v.y = __shfl_sync(0xffffffffu, sum, qb + 1);
    v.z = __shfl_sync(0xffffffffu, sum, qb + 2);
    v.w = __shfl_sync(0xffffffffu, sum, qb + 3);
    if (lane < NV) {
      const float4 bias = *reinterpret_cast<const float4*>(w + (size_t)4 * G * (Cin * K + Cin2) + 4 * g);
      const int c = (int)rank * CoutL + 4 * g;       // first GEMM column of the quad (CoutL == 2: 2 valid columns)
      int oc = c, d = 0;
      if (upsample) {
        d = c / Co;
        oc = c - d * Co;
      }
      const int ol = upsample ? 2 * (l0 + l) + d : l0 + l;
      v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
      if (CoutL >= 4) {
        float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (res_mode != 0) rr = *reinterpret_cast<const float4*>(res + ol * Co + oc);
        if (res_mode == 1) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }       // relu(conv + skip)
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        if (res_mode == 2) { v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }       // relu(convT) + skip
        c2cl_send<4>(out + ol * Co + oc, abar, v, 2u * (unsigned)q, 2u);
      } else {                                       // the 16-column first layer: 2 columns per rank, no residual
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
        c2cl_send<2>(out + ol * Co + oc, abar, v, 2u * (unsigned)q, 2u);
      }
    }
    __syncwarp();                                    // the scratch is reused by this warp's next task
  }
  c2cl_wait(abar, parity);                           // all CoutG x L values of this layer have landed in THIS CTA
}

// in [L][C] -> out [L / 2][C]
__device__ __forceinline__ void c2cl_pool(const float* __restrict__ in, int C, int L, float* __restrict__ out) {
  for (int i = threadIdx.x; i < C * (L / 2); i += C2CL_THREADS) {
    const int l = i / C, c = i - l * C;
    out[i] = fmaxf(in[(2 * l) * C + c], in[(2 * l + 1) * C + c]);
  }
  __syncthreads();
}

// K2 spread over the cluster and over the views: rank r samples the items (z, joint group) i = r, r + 8, ...; a lane group
// (item, view) loads its own taps, then the views' contributions are chained from lane to lane in view order - the same
// fma chain, bit for bit, as the serial loop of c2c_sample_column - and the last view's lane sends the 4 joints' values to
// every CTA's copy of the input columns dst[z][pitch] (counted on `xbar`).
__device__ __forceinline__ void c2cl_sample_column(const FvpPropArgs& a, const FvpSeq& s_seq, int flat, int b, float* __restrict__ dst,
                                                   int pitch, unsigned rank, uint64_t* xbar) {
  const FvpGeom& g = a.g;
  const int Z = g.Z, V = g.V, n_items = Z * g.JG;
  const int cx = flat / g.Y, cy = flat - cx * g.Y;
  const float wx = g.coarse_axes[cx], wy = g.coarse_axes[g.X + cy];
  const float fV = (float)V, rV = 1.0f / fV;
  const int row4 = g.proj.WP * g.JG, px4 = g.JG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipw = 32 / V;                              // items per warp pass
  const int li = lane / V, v = lane - li * V;
  const int mine = (n_items - (int)rank + C2CL_CLUSTER - 1) / C2CL_CLUSTER;       // items of this rank
  for (int base = warp * ipw; base < mine; base += (C2CL_THREADS / 32) * ipw) {
    const int k = base + li;
    const bool on = li < ipw && k < mine;
    const int i = (int)rank + C2CL_CLUSTER * k;
    float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta, tc = ta, td = ta;
    FvpTaps t = {0, 0.f, 0.f, 0.f, 0.f};
    const int z = on ? i / g.JG : 0, s = on ? i - z * g.JG : 0;
    if (on) {
      const float wz = g.coarse_axes[g.X + g.Y + z];
      float ix, iy;
      fvp_project(s_seq.cam[v], s_seq.A, g.proj, wx, wy, wz, ix, iy);
      t = fvp_taps(g.proj, ix, iy);
      const float4* p = (const float4*)a.hm_cl + ((size_t)b * V + v) * g.view_stride4 + s + t.off;
      ta = __ldg(p); tb = __ldg(p + px4); tc = __ldg(p + row4); td = __ldg(p + row4 + px4);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int vv = 0; vv < V; ++vv) {                   // after step vv the lane of view vv holds the sum over views <= vv
      if (vv > 0) {
        const int srcl = li * V + vv - 1;
        acc.x = __shfl_sync(0xffffffffu, acc.x, srcl); acc.y = __shfl_sync(0xffffffffu, acc.y, srcl);
        acc.z = __shfl_sync(0xffffffffu, acc.z, srcl); acc.w = __shfl_sync(0xffffffffu, acc.w, srcl);
      }
      acc.x = fmaf(td.x, t.w11, fmaf(tc.x, t.w10, fmaf(tb.x, t.w01, fmaf(ta.x, t.w00, acc.x))));
      acc.y = fmaf(td.y, t.w11, fmaf(tc.y, t.w10, fmaf(tb.y, t.w01, fmaf(ta.y, t.w00, acc.y))));
      acc.z = fmaf(td.z, t.w11, fmaf(tc.z, t.w10, fmaf(tb.z, t.w01, fmaf(ta.z, t.w00, acc.z))));
      acc.w = fmaf(td.w, t.w11, fmaf(tc.w, t.w10, fmaf(tb.w, t.w01, fmaf(ta.w, t.w00, acc.w))));
    }
    if (on && v == V - 1) c2cl_send<4>(dst + z * pitch + 4 * s, xbar, fvp_mean_clamp4(acc, fV, rV), 0u, (unsigned)C2CL_CLUSTER);
  }
}

__global__ void __cluster_dims__(C2CL_CLUSTER, 1, 1) __launch_bounds__(C2CL_THREADS)
    k_proposals_cluster(FvpPropArgs a, const float* __restrict__ w_head, const float* __restrict__ b_head,
                        const float* __restrict__ w3, int n_cols) {
  extern __shared__ __align__(128) float c2c_smem[];
  __shared__ uint64_t s_wbar[19];                 // weights of layer i have arrived (once per CTA)
  __shared__ uint64_t s_abar[C2CL_NBAR];          // all outputs of layer i (19: the input columns) have arrived (once per column)
  __shared__ int s_woff[19];
  __shared__ FvpSeq s_seq;
  const FvpGeom& g = a.g;
  const int tid = threadIdx.x, J = g.J;
  constexpr int Z = C2CL_Z, L = Z, L2 = Z / 2, L4 = Z / 4;
  const int CJ = 4 * g.JG;                        // channel pitch of the input columns
  const unsigned rank = blockIdx.x % C2CL_CLUSTER;
  const int cluster = blockIdx.x / C2CL_CLUSTER, n_clusters = gridDim.x / C2CL_CLUSTER;
  const int slice = c2cl_slice_floats(J);
  float* s_w = c2c_smem;                          // this rank's weight slice of the whole net
  float* s_buf = s_w + slice;                     // 6 activation buffers (full copies in every CTA)
  float* s_x0 = s_buf + 6 * C2CL_BUF;             // [Z][CJ <= 24] input columns
  float* s_scr = s_x0 + C2CL_X0;                  // split-K scratch of the 4 warps
  float* s_out = s_scr + C2CL_SCRATCH;            // [Z] 1-D heat map (rank 0)
  float *B0 = s_buf, *B1 = s_buf + C2CL_BUF, *B2 = s_buf + 2 * C2CL_BUF, *B3 = s_buf + 3 * C2CL_BUF, *B4 = s_buf + 4 * C2CL_BUF,
        *B5 = s_buf + 5 * C2CL_BUF;

  // ---- the weight slice: one bulk copy per layer, each on its own barrier, issued in consumption order --------------------
  if (tid == 0) {
    int off = 0;
    for (int i = 0; i < 19; ++i) {
      s_woff[i] = off;
      off += c2cl_layer_floats(i, J);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(c2c_s32(s_wbar + i)));
    }
    for (int i = 0; i < C2CL_NBAR; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(c2c_s32(s_abar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const float* src = w3 + (size_t)rank * slice;
    for (int i = 0; i < 19; ++i) {
      const unsigned bytes = (unsigned)c2cl_layer_floats(i, J) * 4u;
      c2cl_expect(s_wbar + i, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                       c2c_s32(s_w + s_woff[i])), "l"(src + s_woff[i]), "r"(bytes), "r"(c2c_s32(s_wbar + i)) : "memory");
    }
  }
  // every CTA of the cluster must have initialised its barriers before anyone stores into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");

  unsigned parity = 0;
  for (int slot = cluster; slot < n_cols; slot += n_clusters, parity ^= 1u) {
    const int b = slot / g.P;
    int flat = 0, seq = 0;
    float c2 = 0.f, sz_w = 0.f, sz_h = 0.f;       // proposal inputs of thread 0 / rank 0, fetched under the network
    const float w_h = rank == 0 ? __ldg(w_head + (tid & 31) * 4) : 0.f, b_h = rank == 0 ? __ldg(b_head) : 0.f;
    if (tid == 0) {                                // this column's arrival counts: CoutG x L values of 4 bytes per layer
      for (int i = 0; i < 19; ++i) {              // (128-column layers run at Z / 4 positions, 64 at Z / 2, the others at Z)
        const int cg = c2c_shape(i, J).CoutG;
        c2cl_expect(s_abar + i, (unsigned)(cg * (cg == 128 ? Z / 4 : cg == 64 ? Z / 2 : Z) * 4));
      }
      if (a.mode == 0) c2cl_expect(s_abar + 19, (unsigned)(Z * g.JG * 4 * 4));
    }
    if (a.mode == 0) {
      flat = a.flat[slot];
      seq = a.frame_seq[b];
      const int* src = (const int*)(g.seqs + seq);
      int* dst = (int*)&s_seq;
      for (int i = tid; i < (int)(sizeof(FvpSeq) / 4); i += C2CL_THREADS) dst[i] = src[i];
      if (rank == 0 && tid == 0) {
        c2 = a.conf2d[slot];
        const float* sz = a.size + (size_t)b * a.size_img_stride + flat;
        sz_w = sz[0];
        sz_h = sz[(size_t)g.X * g.Y];
      }
      __syncthreads();
      c2cl_sample_column(a, s_seq, flat, b, s_x0, CJ, rank, s_abar + 19);
      c2cl_wait(s_abar + 19, parity);
    } else {
      for (int i = tid; i < J * Z; i += C2CL_THREADS) s_x0[(i % Z) * CJ + i / Z] = a.cols_in[(size_t)slot * J * Z + i];
      __syncthreads();
    }
    if (a.cols_out && rank == 0)
      for (int i = tid; i < J * Z; i += C2CL_THREADS) a.cols_out[(size_t)slot * J * Z + i] = s_x0[(i % Z) * CJ + i / Z];

#define CV(Lx, Kx, i, inp, cin, inp2, cin2, outp, cg, resp, rm, up)                                                                      \
  c2cl_conv<Lx, Kx>(inp, (i) == 0 ? CJ : (cin), cin, inp2, cin2, s_w + s_woff[i], s_wbar + i, s_abar + i, parity, outp, cg, resp, rm, up, \
                    rank, s_scr)
    CV(L, 7, 0, s_x0, J, nullptr, 0, B0, 16, nullptr, 0, false);
    CV(L, 3, 1, B0, 16, nullptr, 0, B1, 32, nullptr, 0, false);
    CV(L, 3, 2, B1, 32, B0, 16, B2, 32, nullptr, 0, false);
    CV(L, 3, 3, B2, 32, nullptr, 0, B0, 32, nullptr, 0, false);
    CV(L, 3, 4, B0, 32, nullptr, 0, B3, 32, B2, 1, false);           // skip1
    c2cl_pool(B2, 32, L, B0);
    CV(L2, 3, 5, B0, 32, nullptr, 0, B1, 64, nullptr, 0, false);
    CV(L2, 3, 6, B1, 64, B0, 32, B4, 64, nullptr, 0, false);         // e1
    CV(L2, 3, 7, B4, 64, nullptr, 0, B0, 64, nullptr, 0, false);
    CV(L2, 3, 8, B0, 64, nullptr, 0, B5, 64, B4, 1, false);          // skip2
    c2cl_pool(B4, 64, L2, B0);
    CV(L4, 3, 9, B0, 64, nullptr, 0, B1, 128, nullptr, 0, false);
    CV(L4, 3, 10, B1, 128, B0, 64, B2, 128, nullptr, 0, false);      // e2
    CV(L4, 3, 11, B2, 128, nullptr, 0, B0, 128, nullptr, 0, false);
    CV(L4, 3, 12, B0, 128, nullptr, 0, B1, 128, B2, 1, false);       // m
    CV(L4, 3, 13, B1, 128, nullptr, 0, B0, 128, nullptr, 0, false);
    CV(L4, 3, 14, B0, 128, nullptr, 0, B2, 128, B1, 1, false);       // d2
    CV(L4, 1, 15, B2, 128, nullptr, 0, B0, 128, B5, 2, true);        // u2 = relu(convT)+skip2
    CV(L2, 3, 16, B0, 64, nullptr, 0, B1, 64, nullptr, 0, false);
    CV(L2, 3, 17, B1, 64, nullptr, 0, B2, 64, B0, 1, false);         // d1
    CV(L2, 1, 18, B2, 64, nullptr, 0, B4, 64, B3, 2, true);          // u1 = relu(convT)+skip1 (into B4, see the hazard note)
#undef CV
    // head: 1x1 32 -> 4, only channel 0 is the 1-D heat map; rank 0 alone, weights [ci][4] straight from L2
    if (rank == 0) {
      const int warp = tid >> 5, lane = tid & 31;
      for (int t = warp; t < L / C2CL_TILE; t += C2CL_THREADS / 32) {
        const float wv = w_h;
        float acc[C2CL_TILE];
#pragma unroll
        for (int l = 0; l < C2CL_TILE; ++l) acc[l] = wv * B4[(t * C2CL_TILE + l) * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int l = 0; l < C2CL_TILE; ++l) acc[l] += __shfl_xor_sync(0xffffffffu, acc[l], o);
        if (lane == 0) {
#pragma unroll
          for (int l = 0; l < C2CL_TILE; ++l) s_out[t * C2CL_TILE + l] = acc[l] + b_h;
        }
      }
      __syncthreads();
      if (a.hm1d_out)
        for (int i = tid; i < Z; i += C2CL_THREADS) a.hm1d_out[(size_t)slot * Z + i] = s_out[i];
      if (a.mode == 0 && tid == 0) c2c_emit_proposal(a, slot, seq, flat, s_out, c2, sz_w, sz_h);
      __syncthreads();                             // s_out is rewritten by the next column's head
    }
  }
  // nobody leaves while a peer could still address its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__global__ void k_people_from_centers(FvpPropArgs a, const float* __restrict__ centers, int n) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  float c7[7];
  for (int i = 0; i < 7; ++i) c7[i] = centers[(size_t)slot * 7 + i];
  FvpPerson pd;
  fvp_make_person(a, c7, a.frame_seq[slot / a.g.P], pd);
  a.people[slot] = pd;
  if (a.img_valid)
    for (int q = 0; q < 3; ++q) a.img_valid[q * a.n_slots + slot] = pd.valid;
}

// Per-device setup (fvp_create): dynamic shared memory opt-ins of the NMS kernel (the whole X*Y map) and the proposal kernel.
// A grid whose NMS map does not fit the 227 KB per-CTA limit is refused here, at create time, with a clear error.
cudaError_t fvp_proposal_init_device(int X, int Y) {
  const size_t nms = fvp_nms_smem_bytes(X, Y);
  if (nms + 1024 > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaFuncAttributes fa;                         // several contexts may share the device: never lower an earlier opt-in
  cudaError_t e = cudaFuncGetAttributes(&fa, k_nms_topk);
  if (e != cudaSuccess) return e;
  if (nms > 48 * 1024 && nms > (size_t)fa.maxDynamicSharedSizeBytes)
    e = cudaFuncSetAttribute(k_nms_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms);
  if (e != cudaSuccess) return e;
  const size_t c2c = c2c_smem_bytes(20) > c2c_smem_bytes(40) ? c2c_smem_bytes(20) : c2c_smem_bytes(40);
  e = cudaFuncSetAttribute(k_proposals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2c);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_proposals_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2cl_smem_bytes(24));
}

// Per-rank weight slices of the cluster kernel from the HOST ci-major copies ([ci][tap][CoutG], then [ci2][CoutG]):
// out[rank][layer] = [g][ci][tap][4] + [g][ci2][4] + [g][4] biases (see the layout note at k_proposals_cluster).  Returns the floats written (8 x slice), 0 if J is unsupported.
int fvp_c2c_max_clusters(int J) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 / C2CL_CLUSTER * C2CL_CLUSTER);
  cfg.blockDim = dim3(C2CL_THREADS);
  cfg.dynamicSmemBytes = c2cl_smem_bytes(J <= 24 ? J : 24);
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = C2CL_CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_proposals_cluster, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 16;
  }
  return n;
}
size_t fvp_c2c_cluster_floats(int J) { return J <= 24 ? (size_t)C2CL_CLUSTER * c2cl_slice_floats(J) : 0; }
void fvp_c2c_pack_cluster(const float* const w2_host[20], const float* const b_host[20], int J, float* out) {
  const int slice = c2cl_slice_floats(J);
  for (int r = 0; r < C2CL_CLUSTER; ++r) {
    float* dst = out + (size_t)r * slice;           // zero filled by the caller: the padding channels stay zero
    for (int i = 0; i < 19; ++i) {
      const C2CLayerShape sh = c2c_shape(i, J);
      const int CoutL = sh.CoutG / C2CL_CLUSTER, G = (CoutL + 3) / 4;
      const float* src = w2_host[i];
      float* ds = dst + (size_t)4 * G * sh.Cin * sh.K;
      float* db = ds + (size_t)4 * G * sh.Cin2;
      for (int cl = 0; cl < CoutL; ++cl) {
        const int g = cl >> 2, q = cl & 3, co = r * CoutL + cl;
        for (int ci = 0; ci < sh.Cin; ++ci)
          for (int tp = 0; tp < sh.K; ++tp)
            dst[((size_t)(g * sh.Cin + ci) * sh.K + tp) * 4 + q] = src[(size_t)(ci * sh.K + tp) * sh.CoutG + co];
        for (int ci = 0; ci < sh.Cin2; ++ci) ds[(size_t)(g * sh.Cin2 + ci) * 4 + q] = src[(size_t)(sh.Cin * sh.K + ci) * sh.CoutG + co];
        db[cl] = b_host[i][co];
      }
      dst += c2cl_layer_floats(i, J);
    }
  }
}

// Chunk table of the network-wide weight ring: (layer, part, chunk) in the order c2c_forward2 consumes them.  Written to
// `h_plan` (host) - fvp_pack_params uploads it next to the weights and passes the device copy back in FvpC2CW::plan.
size_t fvp_c2c_plan_bytes() { return sizeof(C2CPlan); }
int fvp_c2c_build_plan(const float* const w2[20], int J, void* h_plan) {
  C2CPlan& p = *reinterpret_cast<C2CPlan*>(h_plan);
  p.n = 0;
  for (int i = 0; i < 20; ++i) {
    const C2CLayerShape sh = c2c_shape(i, J);
    const float* base = w2[i];
    for (int part = 0; part < 2; ++part) {
      const int C = part == 0 ? sh.Cin : sh.Cin2, taps = part == 0 ? sh.K : 1;
      if (C == 0) continue;
      const float* src = part == 0 ? base : base + (size_t)sh.Cin * sh.K * sh.CoutG;
      const int cpc = c2c_cpc(C, taps, sh.CoutG);
      for (int c_lo = 0; c_lo < C; c_lo += cpc) {
        if (p.n >= C2C_MAX_CHUNKS) return -1;
        const int nc = C - c_lo < cpc ? C - c_lo : cpc;
        p.chunk[p.n].src = src + (size_t)c_lo * taps * sh.CoutG;
        p.chunk[p.n].nfloats = nc * taps * sh.CoutG;
        p.chunk[p.n].pad = 0;
        if ((p.chunk[p.n].nfloats & 3) || ((uintptr_t)p.chunk[p.n].src & 15)) return -2;     // TMA bulk: 16-byte granules
        ++p.n;
      }
    }
  }
  return p.n;
}

void fvp_launch_proposals(const FvpPropArgs& a, const FvpC2CW& w, int n, int prefer_latency, cudaStream_t st) {
  C2CNet net;
  for (int i = 0; i < 20; ++i) {
    net.l[i].w = w.w[i];
    net.l[i].b = w.b[i];
  }
  C2CNet2 net2;
  for (int i = 0; i < 20; ++i) net2.b[i] = w.b[i];
  if (prefer_latency && w.w3 != nullptr && a.g.Z == C2CL_Z && n <= w.max_clusters) {     // one 8-CTA cluster per column
    k_proposals_cluster<<<n * C2CL_CLUSTER, C2CL_THREADS, c2cl_smem_bytes(a.g.J), st>>>(a, w.w2[19], w.b[19], w.w3, n);
    return;
  }
  const size_t smem = c2c_smem_bytes(a.g.Z);
  k_proposals<<<n, C2C_THREADS, smem, st>>>(a, net, net2, reinterpret_cast<const C2CPlan*>(w.plan), c2c_nst(a.g.Z));
}

void fvp_launch_people_from_centers(const FvpPropArgs& a, const float* d_centers, int n, cudaStream_t st) {
  k_people_from_centers<<<fvp_cdiv(n, 64), 64, 0, st>>>(a, d_centers, n);
}
