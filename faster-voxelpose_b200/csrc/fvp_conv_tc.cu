// fvp_conv_tc.cu - tcgen05 / TMEM implicit-GEMM convolution for the CenterNet / P2PNet trunks
// (cnns_2d.py:12-178), same fused epilogues and NHWC fp32 interface as fvp_conv.cu (k_conv_nhwc).
//
// Precision: two error-compensated operand splits (template parameter F16):
//   F16 = true  (default): x = hi + lo * 2^-11 with hi = fp16(x), lo = fp16((x - hi) * 2^11); kind::f16 MMAs with K = 16,
//                D1 += A_hi*B_hi and D2 += A_hi*B_lo + A_lo*B_hi in two fp32 TMEM accumulators, D = D1 + 2^-11 * D2.
//                22 significant bits per operand like 3xTF32 but HALF the MMA instructions (the tensor unit spends
//                ~48 cycles of fixed overhead per instruction with cta_group::1); operands must stay
//                inside the fp16 range (|x| < 65504 - activations here are O(1)).
//   F16 = false: "3xTF32" (below).
// Precision (tf32 variant): "3xTF32" error-compensated split.  Every fp32 operand x is split on the fly into
// hi = tf32(x) (round-to-nearest) and lo = tf32(x - hi); D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi accumulates
// in fp32 in tensor memory.  A single kind::tf32 pass moves joints by 1.33 mm; the split keeps
// the convolution at ~1e-6 relative of the exact-fp32 kernel (tests/test_gpu_parity.py compares both).
//
// GEMM view per CTA: M = 128 output pixels (tile 16 rows x 8 cols), N = N_TILE <= 128 output channels,
// K = taps x Cin (+ Cin2 of the fused 1x1 skip conv).
//   * A (activations): the (16+k-1)x(8+k-1) input halo of one K-block (32 or 16 channels) is staged ONCE in shared memory
//     as pixel rows in the canonical K-major swizzled UMMA layout, hi plane then lo plane, with the halo row pitch padded
//     to 16 pixels.  Because the tile is 8 pixels wide, the operand of tap (dy,dx) is the same buffer with its start
//     address advanced by (dy*16+dx) rows: nine taps = nine descriptors, no im2col copies.
//   * B (weights): pre-split hi/lo and pre-tiled on the host into the exact shared-memory image of each
//     (K-block, tap, N-tile): whole image resident, one N tile resident, or streamed per tap by TMA bulk copies.
//   * D: 128 lanes x N_TILE fp32 columns of TMEM (two buffers, D1|D2 side by side in the fp16 modes); epilogue =
//     tcgen05.ld 32x32b, D1 + 2^-11 D2, bias/residual/ReLU, NHWC store (or pixel-shuffle for ConvTranspose k2s2, or planar
//     store for the final layers).
// Warp roles (14 warps, see the kernel): 8 loader warps | MMA issuer + TMEM owner | weight TMA producer | 4 epilogue warps,
// connected by mbarrier full/empty rings; MMA completion frees a slot through tcgen05.commit.
// DESIGN.md 4.1 has the measured hardware facts this layout rests on and what bounds the kernel.
#include <cuda.h>          // CUtensorMap + enums only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fvp_kernels.h"

// Range guard of the fp16 hi/lo engine (per device, set by fvp_create): every stored activation is the next layer's
// operand and must convert to a finite fp16 "hi" part (network inputs are back-projected planes clamped to [0,1]).  An output with |x| >= 65504 (or NaN) raises bit 0 of the status
// word in mapped pinned host memory; the host entry points turn it into FVP_E_RANGE after their next synchronise.
__device__ int* g_tc_status = nullptr;

namespace {

constexpr int TC_TH = 16, TC_TW = 8;       // output tile (pixels)
constexpr int TC_LOADERS = 256;            // warps 0..7: A staging (16 loader warps in the one-CTA variant measured 13-32 % slower
                                           // on every layer, profiles/r02_conv_experiments.txt)
constexpr int TC_LW = TC_LOADERS / 32;     // warp 8 = MMA issuer / TMEM owner, warp 9 = weight TMA producer, warps 10-13 = epilogue
constexpr int TC_THREADS = TC_LOADERS + 64 + 128;
constexpr int TC_EPI = 128;
constexpr int TC_MAX_A = 2, TC_MAX_B = 8;  // ring depths are chosen per launch (a_stages, b_stages)
constexpr int TC_MAX_COUT = 256;           // bias staged in static shared memory (largest layer: ConvTranspose 128->4x64)
constexpr int TC_STATIC_SMEM = 1280;       // s_bar + s_tmem + s_bias, as reported by ptxas (checked in fvp_launch_conv_tc)
constexpr int TC_DYN_SMEM_MAX = 227 * 1024 - TC_STATIC_SMEM;   // opt-in limit per CTA minus the static part

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared by the TMA engine (SASS UBLKCP), completion signalled on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 4-D tiled TMA tensor load global -> shared (SASS UTMALDG): box = (channels of one K-block, halo width padded to the
// row pitch, halo height, 1 image), out-of-bounds elements arrive as zeros (= the convolution's zero padding), the
// 16-byte chunks land swizzled exactly as the UMMA descriptors of this kernel read them.
__device__ __forceinline__ void tma_tensor4d_g2s(void* dst_smem, const CUtensorMap* map, int c0, int x, int y, int img, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
          smem_u32(dst_smem)),
      "l"(map), "r"(c0), "r"(x), "r"(y), "r"(img), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 28)) {        // watchdog: a protocol bug must abort the kernel, never hang the GPU
      printf("k_conv_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// One lane of a CONVERGED warp (the MMA warp enters its role through a warp-uniform branch, see below): ptxas then
// emits ELECT + a predicated UTCHMMA instead of the per-instruction "for every active lane" loop it generates around
// uniform-datapath instructions in divergent code (6 extra instructions and ~40 cycles per MMA, measured).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 B (32 tf32), 16-B
// chunks XOR-swizzled with the row index mod 8.
//   [0,14) start>>4, [16,30) leading byte offset>>4 (unused for swizzled K-major: 1), [32,46) stride byte offset>>4
//   (stride between 8-row groups), [46,48) version = 1, [49,52) base offset (left 0), [61,64) layout type = 2.
// Measured on B200 (tools/conv_debug.py): the hardware applies the 128-B swizzle to ABSOLUTE shared-memory address
// bits (chunk ^= (addr >> 7) & 7), so a tap-shifted operand that starts in the middle of a 1024-B atom needs no
// base offset as long as the data was stored with the same absolute-address rule (setting base_offset = dx breaks it).
// (Round-1 note: the un-swizzled "interleave" layout computes correctly but the tensor core then fetches one 16-B row
//  per cycle: ~(128+N)*2 cycles per MMA, 13x below peak - measured in profiles/r01_launches_tc_noswizzle.csv.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);     // layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N>>3, M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A = B = F16 (format 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One tap of the fp16 engine as a single instruction group: A_hi x [B_hi;B_lo] -> D1|D2 (width 2n), A_lo x B_hi -> D2
// (width n), for one (KSTEPS = 1) or two (KSTEPS = 2) K = 16 steps.  Descriptors advance by 2 (32 B) per K step.
template <int KSTEPS>
__device__ __forceinline__ void umma_tap_f16(uint32_t d1, uint32_t d2, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint32_t idesc2,
                                             uint32_t idesc, uint32_t accumulate) {
  if constexpr (KSTEPS == 2) {
    asm volatile(
        "{\n\t.reg .pred p, pt;\n\t.reg .b64 a2, al2, b2;\n\t"
        "setp.ne.b32 p, %7, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 a2, %2, 2;\n\tadd.u64 al2, %3, 2;\n\tadd.u64 b2, %4, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %5, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %5, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %6, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], al2, b2, %6, pt;\n\t}\n" ::"r"(d1),
        "r"(d2), "l"(a_hi), "l"(a_lo), "l"(b_hi), "r"(idesc2), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, pt;\n\t"
        "setp.ne.b32 p, %7, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %5, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %6, pt;\n\t}\n" ::"r"(d1),
        "r"(d2), "l"(a_hi), "l"(a_lo), "l"(b_hi), "r"(idesc2), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// All taps of one K-block of the fp16 engine with resident weights, fully unrolled: the tap offsets into the halo are
// compile-time immediates on one base descriptor, the weight descriptor advances by a constant stride - no per-tap
// scalar work is left on the issuing lane (measured: ~130 cycles of per-tap overhead in the generic loop below).
template <int ROWB, int K, int KSTEPS>
__device__ __forceinline__ void umma_kblock_f16(uint32_t d1, uint32_t d2, uint64_t ad0, uint32_t a_lo16, uint64_t bd, uint32_t b_stride16,
                                                uint32_t idesc2, uint32_t idesc, uint32_t accumulate) {
  constexpr int HWP = K == 1 ? 8 : 16;
#pragma unroll
  for (int dy = 0; dy < K; ++dy) {
#pragma unroll
    for (int dx = 0; dx < K; ++dx) {
      const uint64_t ah = ad0 + (uint32_t)((dy * HWP + dx) * (ROWB >> 4));
      umma_tap_f16<KSTEPS>(d1, d2, ah, ah + a_lo16, bd, idesc2, idesc, (dy | dx) ? 1u : accumulate);
      bd += b_stride16;
    }
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// issue only: the destination registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
template <int CW>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t* r) {
  if constexpr (CW == 16) tmem_ld16_issue(taddr, r);
  else tmem_ld8_issue(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// NOTE: ptxas's register allocation of the 72-register (two CTAs per SM) variants is sensitive to this struct: adding one int
// field raised the spills of k_conv_tc<1,2> from 16 B to 116 B and made the 3x3 32->32 layer 45 % slower (local memory thrashes
// the 28 KB of L1 left beside 2 x 111 KB of shared memory).  tests/test_host_logic.py pins the spill sizes from the build log.
struct TcArgs {
  FvpConvArgs c;
  const float* wtc;      // tiled hi/lo weights (see pack_tc in fvp_params.cu)
  int n_tile;            // GEMM N of this launch (multiple of 16, <= 128)
  int n_tiles;           // CoutPad / n_tile
  uint32_t a_stage_bytes, b_stage_bytes, tmem_cols, acc_stride;
  int a_stages, b_stages;
  int resident;          // 1: the whole weight image is loaded once per CTA (b_stage_bytes = its size); 2: only the blocks of
                         // ONE N tile (b_stage_bytes / blk_bytes of them) - the grid is a multiple of n_tiles, so CTA b only ever sees N tile b % n_tiles
  uint32_t blk_bytes;    // bytes of one (K-block, tap, N-tile) weight block = n_tile*128*2
  int tiles_x, tiles_per_img, total_items;   // work item = (image, tile, N tile)
  uint32_t m_ntiles, m_tpi, m_tx;            // ceil(2^32 / d) of the three divisors of tc_decode (0: use the division)
  unsigned long long* prof;                  // debug: per-role wait / busy cycle counters (NULL in production)
};

// TMA descriptors of a launch whose inputs are split tensors (fp16 hi / lo planes): main input and fused-skip input
struct alignas(64) TcMaps {
  CUtensorMap in_hi, in_lo, in2_hi, in2_lo;
};

// debug instrumentation: add the cycles spent in `stmt` to counter `slot` when profiling is on
#define TC_TIMED(slot, stmt)                                   \
  do {                                                         \
    if (t.prof) {                                              \
      const long long c0_ = clock64();                         \
      stmt;                                                    \
      pc[slot] += clock64() - c0_;                             \
    } else {                                                   \
      stmt;                                                    \
    }                                                          \
  } while (0)

struct TcItem {
  int img, y0, x0, nt;
};
// n / d through one multiply-high with m = ceil(2^32 / d): exact while n * d < 2^32 (checked on the host, m = 0 otherwise)
__device__ __forceinline__ int tc_div(int n, int d, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n / d; }
__device__ __forceinline__ bool tc_decode(const TcArgs& t, int item, TcItem& w) {
  const int r = t.n_tiles == 1 ? item : tc_div(item, t.n_tiles, t.m_ntiles);
  w.nt = item - r * t.n_tiles;
  w.img = tc_div(r, t.tiles_per_img, t.m_tpi);
  const int tile = r - w.img * t.tiles_per_img;
  const int ty = tc_div(tile, t.tiles_x, t.m_tx);
  w.y0 = ty * TC_TH;
  w.x0 = (tile - ty * t.tiles_x) * TC_TW;
  return !(t.c.valid && !t.c.valid[w.img]);
}

// Persistent CTA (one or two per SM), 14 warps:
//   warps 0-7    A staging   : halo of the next K-block -> hi/lo planes                   (a_full / a_empty ring)
//   warp  8      MMA issue   : one elected lane issues every tcgen05.mma; owns the TMEM allocation
//   warp  9      B producer  : one lane issues the weight TMA bulk copies                 (b_full / b_empty ring)
//   warps 10-13  epilogue    : tcgen05.ld of the finished accumulator, bias/residual/ReLU, stores
// Two accumulator buffers in TMEM (acc_full / acc_empty) let the MMAs of tile i+1 run under the epilogue of
// tile i, and the staging of tile i+1 under the MMAs of tile i.
// MODE 0: 3xTF32, 32-channel K-blocks (128-B rows, SWIZZLE_128B); MODE 1: fp16 split, 32-channel K-blocks (64-B rows,
// SWIZZLE_64B); MODE 2: fp16 split, 16-channel K-blocks (32-B rows, SWIZZLE_32B) for layers with <= 16 input channels.
// OCC = CTAs per SM the kernel is compiled for: 2 caps the registers at 72 so that two persistent CTAs (each with its own
// loaders / MMA issuer / epilogue) share an SM when the launch needs <= ~110 KB of shared memory and <= 256 TMEM columns -
// every role of this kernel is latency-bound per item, a second CTA fills the idle issue slots and tensor-pipe gaps.
// TMA = true: the input (and fused-skip input) are split tensors fetched by TMA tensor loads from the producer lane - the 8
// loader warps do not exist (192 threads: MMA warp, producer warp, 4 epilogue warps).
template <int MODE, int OCC, bool TMA>
__global__ void __launch_bounds__(TMA ? 192 : TC_THREADS, OCC) k_conv_tc(TcArgs t, const __grid_constant__ TcMaps maps) {
  constexpr bool F16 = MODE != 0;
  static_assert(!TMA || F16, "split tensors are the operand format of the fp16 engine");
  constexpr int LW = TMA ? 0 : TC_LW;                               // warps in front of the MMA warp (the loaders)
  constexpr int NTHREADS = LW * 32 + 64 + TC_EPI;
  // (Measured, profiles/r02_conv_ab.txt: compiling the streamed-weight path out of the two-CTA instantiations - which can
  // never stream, streaming fills the shared memory - lowers their spill counts but made the 7x7 layer 11 % SLOWER; the
  // 72-register variants are a register-allocation lottery, only measurements count.  The path stays in every variant.)
  constexpr bool CAN_STREAM = true;
  constexpr int ROWB = MODE == 0 ? 128 : (MODE == 1 ? 64 : 32);   // bytes of one pixel row of a K-block in shared memory
  constexpr int CB = MODE == 2 ? 16 : 32;                          // channels per K-block
  constexpr uint32_t LAYOUT = MODE == 0 ? 2u : (MODE == 1 ? 4u : 6u);   // SWIZZLE_128B / 64B / 32B
  constexpr int PH_SHIFT = MODE == 0 ? 0 : (MODE == 1 ? 1 : 2);    // swizzle phase = (row index >> PH_SHIFT) & (chunks - 1)
  extern __shared__ __align__(128) uint8_t tc_smem[];
  __shared__ uint64_t s_bar[2 * TC_MAX_A + 2 * TC_MAX_B + 4];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[TC_MAX_COUT];              // bias of every output channel
  const FvpConvArgs& a = t.c;
  // per-N-tile weight residency needs > 113 KB of shared memory, i.e. it only exists in the one-CTA-per-SM variant (keeping it
  // out of the 72-register variant avoids spills there: local memory thrashes the 28 KB of L1 left beside 2 x 111 KB of smem)
  const bool per_nt = OCC == 1 && t.resident == 2;

  const int A_ST = t.a_stages, B_ST = t.b_stages;
  uint8_t* sA = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);   // swizzle atoms need 1024-B alignment
  uint8_t* sB = sA + A_ST * t.a_stage_bytes;                       // [B_ST][b_stage_bytes] (hi then lo)
  uint64_t* a_full = s_bar;
  uint64_t* a_empty = s_bar + TC_MAX_A;
  uint64_t* b_full = s_bar + 2 * TC_MAX_A;
  uint64_t* b_empty = b_full + TC_MAX_B;
  uint64_t* acc_full = b_empty + TC_MAX_B;                         // [2]
  uint64_t* acc_empty = acc_full + 2;                              // [2]

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform role index
  // Programmatic dependent launch: let the next kernel of the stream start its own prologue as soon as SMs are free,
  // and run OUR prologue (barriers, TMEM allocation, resident weight fetch) under the tail of the previous kernel.
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if (tid == 0) {
    for (int i = 0; i < A_ST; ++i) { mbar_init(a_full + i, TMA ? 1 : TC_LOADERS); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < B_ST; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, TC_EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < a.CoutP; i += NTHREADS) s_bias[i] = a.bias[i];   // weights: independent of the previous kernel
  if (warp == LW) {                                                // TMEM allocation (power of two >= 32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(t.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int nph = a.in2 ? 2 : 1;
  const bool is_producer = tid == LW * 32 + 32;
  if (is_producer && t.resident) {                                // weights do not depend on the previous kernel
    mbar_arrive_expect_tx(b_full, t.b_stage_bytes);
    if (per_nt) {                                         // my N tile's blocks, compacted: block b -> sB + b * blk_bytes
      const uint32_t nt = blockIdx.x % (uint32_t)t.n_tiles;
      const int nblk = (int)(t.b_stage_bytes / t.blk_bytes);     // (K-block, tap) blocks of one N tile
      for (int b = 0; b < nblk; ++b)
        tma_bulk_g2s(sB + (size_t)b * t.blk_bytes, (const uint8_t*)t.wtc + ((size_t)b * t.n_tiles + nt) * t.blk_bytes, t.blk_bytes, b_full);
    } else {
      for (uint32_t off = 0; off < t.b_stage_bytes; off += 32768) {
        const uint32_t n = t.b_stage_bytes - off < 32768 ? t.b_stage_bytes - off : 32768;
        tma_bulk_g2s(sB + off, (const uint8_t*)t.wtc + off, n, b_full);
      }
    }
  }
  asm volatile("griddepcontrol.wait;\n" ::: "memory");           // activations of the previous kernel are now visible

  if (!TMA && warp < TC_LW) {
    // =============================== A staging (256 threads) ========================================
    // Thread (q = tid & 7, p0 = tid >> 3) always handles channel quad q of halo pixels p0, p0+32, p0+64, ...:
    // their halo coordinates, shared-memory destinations and global offsets do not depend on the work item, so they
    // are computed once; per item only the image bounds tests and one base pointer remain.
    constexpr int CH = ROWB / 16;                                  // 16-B chunks per pixel row
    constexpr int PPI = TC_LOADERS / CH;                           // pixels covered per pass of the 256 threads
    constexpr int MAXHALO = MODE == 2 ? 22 * 14 : 18 * 10;         // 7x7 layers always run in MODE 2 (launcher)
    constexpr int EPT = (MAXHALO + PPI - 1) / PPI;                 // elements per thread
    constexpr int NV = F16 ? 2 : 1;                                // float4 loads per element
    const int q = tid % CH, p0 = tid / CH;
    int e_hyx[2][EPT], e_dst[2][EPT];                              // [phase][element]: (hy<<8)|hx ; smem byte offset or -1
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const int K = ph == 0 ? a.ksize : 1;
      const int HW = TC_TW + K - 1, HH = TC_TH + K - 1, HWP = K == 1 ? 8 : 16;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int pix = p0 + PPI * j, hy = pix / HW, hx = pix - hy * HW;
        e_hyx[ph][j] = (hy << 8) | hx;
        // swizzle on absolute address bits: chunk ^= (row address >> 7) & (CH-1); halo rows are 1024-B multiples apart
        const int phase = (hx >> PH_SHIFT) & (CH - 1);
        e_dst[ph][j] = pix < HW * HH ? (hy * HWP + hx) * ROWB + ((q ^ phase) << 4) : -1;
      }
    }
    // The K-blocks of this CTA form one flat sequence (item, phase, 32-channel block).  The global loads of block i+1
    // are issued BEFORE block i is converted and stored, so their L2 latency (the loaders' largest stall, ncu: 38 % of
    // the samples on the first dependent convert) runs under the conversion and the wait for a free stage.
    struct KBlock { const float* img; int y0, x0, c0, ph; };
    int it_item = blockIdx.x, it_ph = 0, it_c0 = 0;
    bool it_open = false;
    TcItem it_w;
    auto advance = [&](KBlock& kb) -> bool {
      while (it_item < t.total_items) {
        if (!it_open) {
          if (!tc_decode(t, it_item, it_w)) { it_item += gridDim.x; continue; }
          it_open = true; it_ph = 0; it_c0 = 0;
        }
        const int Cin = it_ph == 0 ? a.Cin : a.Cin2;
        if (it_c0 >= Cin) {                                         // K-blocks of this phase are done (CinP = round-up)
          it_c0 = 0;
          if (++it_ph >= nph) { it_open = false; it_item += gridDim.x; }
          continue;
        }
        kb.img = (it_ph == 0 ? a.in : a.in2) + (size_t)it_w.img * a.H * a.W * Cin;
        kb.y0 = it_w.y0; kb.x0 = it_w.x0; kb.c0 = it_c0; kb.ph = it_ph;
        it_c0 += CB;
        return true;
      }
      return false;
    };
    auto issue = [&](const KBlock& kb, float4 (&v)[EPT][NV]) {
      const int K = kb.ph == 0 ? a.ksize : 1, Cin = kb.ph == 0 ? a.Cin : a.Cin2, pad = (K - 1) / 2;
      const int c = kb.c0 + q * (F16 ? 8 : 4);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int hyx = kb.ph == 0 ? e_hyx[0][j] : e_hyx[1][j], dst = kb.ph == 0 ? e_dst[0][j] : e_dst[1][j];
        const int gy = kb.y0 + (hyx >> 8) - pad, gx = kb.x0 + (hyx & 255) - pad;
#pragma unroll
        for (int h = 0; h < NV; ++h) v[j][h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dst >= 0 && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float* pp = kb.img + ((size_t)gy * a.W + gx) * Cin + c;
#pragma unroll
          for (int h = 0; h < NV; ++h)
            if (c + 4 * h < Cin) v[j][h] = __ldg((const float4*)(pp + 4 * h));
        }
      }
    };
    int a_it = 0;
    long long pc[2] = {0, 0};
    const long long lt0 = clock64();
    auto stage = [&](const KBlock& kb, const float4 (&v)[EPT][NV]) {
      const int K = kb.ph == 0 ? a.ksize : 1;
      const int HH = TC_TH + K - 1, HWP = K == 1 ? 8 : 16;
      const uint32_t lo_off = (uint32_t)HH * HWP * ROWB;
      const int as = a_it % A_ST;
      if (a_it >= A_ST) TC_TIMED(0, mbar_wait(a_empty + as, ((a_it / A_ST) - 1) & 1));
      uint8_t* hi = sA + (size_t)as * t.a_stage_bytes;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int dst = kb.ph == 0 ? e_dst[0][j] : e_dst[1][j];
        if (dst < 0) continue;
        if constexpr (F16) {
          const float x[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w, v[j][NV - 1].x, v[j][NV - 1].y, v[j][NV - 1].z, v[j][NV - 1].w};
          uint32_t ph_[4], pl_[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {                            // packed converts (F2FP on the ALU pipe, not the
            const __half2 h2 = __floats2half2_rn(x[2 * e], x[2 * e + 1]);   // quarter-rate scalar F2F)
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn((x[2 * e] - hf.x) * 2048.0f, (x[2 * e + 1] - hf.y) * 2048.0f);
            ph_[e] = *reinterpret_cast<const uint32_t*>(&h2);
            pl_[e] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          *(uint4*)(hi + dst) = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
          *(uint4*)(hi + lo_off + dst) = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
        } else {
          const float4 x = v[j][0];
          const float4 h = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
          *(float4*)(hi + dst) = h;
          *(float4*)(hi + lo_off + dst) = make_float4(to_tf32(x.x - h.x), to_tf32(x.y - h.y), to_tf32(x.z - h.z), to_tf32(x.w - h.w));
        }
      }
      fence_proxy_async();
      mbar_arrive(a_full + as);
      ++a_it;
    };
    if constexpr (OCC == 1) {
      float4 va[EPT][NV], vb[EPT][NV];
      KBlock ka, kb;
      bool more = advance(ka);
      if (more) issue(ka, va);
      while (more) {
        const bool more_b = advance(kb);
        if (more_b) issue(kb, vb);
        stage(ka, va);
        if (!more_b) break;
        more = advance(ka);
        if (more) issue(ka, va);
        stage(kb, vb);
      }
    } else {                                                       // 72-register variant: the sibling CTA hides the latency
      float4 va[EPT][NV];
      KBlock ka;
      while (advance(ka)) {
        issue(ka, va);
        stage(ka, va);
      }
    }
    if (t.prof && tid == 0) {
      atomicAdd(t.prof + 0, (unsigned long long)pc[0]);
      atomicAdd(t.prof + 1, (unsigned long long)(clock64() - lt0));
    }
  } else if (tid == LW * 32 + 32) {
    // ============== producer lane: weight TMA bulk copies (+ the activation halos when TMA) ==============
    if (TMA || (CAN_STREAM && !t.resident)) {                     // (resident image: already requested above)
      int b_it = 0, a_it = 0;
      for (int item = blockIdx.x; item < t.total_items; item += gridDim.x) {
        TcItem w;
        if (!tc_decode(t, item, w)) continue;
        const uint8_t* wsrc = (const uint8_t*)t.wtc;
        for (int ph = 0; ph < nph; ++ph) {
          const int K = ph == 0 ? a.ksize : 1, Cin = ph == 0 ? a.Cin : a.Cin2;
          const int CinP = (Cin + CB - 1) / CB * CB;
          for (int c0 = 0; c0 < CinP; c0 += CB) {
            if constexpr (TMA) {
              // halo of this K-block: hi plane then lo plane, one tensor load each.  Order A(kb), B(kb, rows), A(kb+1), ...:
              // the MMA warp needs them in exactly this order, so waiting for a free A stage can never starve it of weights.
              const int HH = TC_TH + K - 1, HWP = K == 1 ? 8 : 16, pad = (K - 1) / 2;
              const uint32_t plane = (uint32_t)HH * HWP * ROWB;
              const int as = a_it % A_ST;
              if (a_it >= A_ST) mbar_wait(a_empty + as, ((a_it / A_ST) - 1) & 1);
              uint8_t* dst = sA + (size_t)as * t.a_stage_bytes;
              mbar_arrive_expect_tx(a_full + as, 2 * plane);
              tma_tensor4d_g2s(dst, ph == 0 ? &maps.in_hi : &maps.in2_hi, c0, w.x0 - pad, w.y0 - pad, w.img, a_full + as);
              tma_tensor4d_g2s(dst + plane, ph == 0 ? &maps.in_lo : &maps.in2_lo, c0, w.x0 - pad, w.y0 - pad, w.img, a_full + as);
              ++a_it;
            }
            if (!(CAN_STREAM && !t.resident)) continue;
            // One ring stage = the K weight blocks of one tap ROW: one barrier wait + one commit per K taps on the issuing
            // lane instead of per tap (measured on the streamed 128-channel layers at 960 images: 229 -> 184 us for
            // 128->128, 128 -> 104 us for 64->128; profiles/r02_conv_experiments.txt).
            for (int dy = 0; dy < K; ++dy) {
              const int bs = b_it % B_ST;
              if (b_it >= B_ST) mbar_wait(b_empty + bs, ((b_it / B_ST) - 1) & 1);
              mbar_arrive_expect_tx(b_full + bs, (uint32_t)K * t.blk_bytes);
              for (int dx = 0; dx < K; ++dx)
                tma_bulk_g2s(sB + (size_t)bs * t.b_stage_bytes + (size_t)dx * t.blk_bytes,
                             wsrc + ((size_t)(dy * K + dx) * t.n_tiles + w.nt) * t.blk_bytes, t.blk_bytes, b_full + bs);
              ++b_it;
            }
            wsrc += (size_t)K * K * t.n_tiles * t.blk_bytes;
          }
        }
      }
    }
  } else if (warp == LW) {
    // ======================= MMA issue (whole warp converged, one elected lane issues) ===============
    const uint32_t idesc = F16 ? umma_idesc_f16(128, t.n_tile) : umma_idesc_tf32(128, t.n_tile);
    const uint32_t idesc2 = umma_idesc_f16(128, 2 * t.n_tile);   // fp16 engine: A_hi x [B_hi ; B_lo]
    int a_it = 0, b_it = 0, it = 0;
    bool b_ready = false;
    long long pc[4] = {0, 0, 0, 0};
    const long long mt0 = clock64();
    for (int item = blockIdx.x; item < t.total_items; item += gridDim.x) {
      TcItem w;
      if (!tc_decode(t, item, w)) continue;
      const int buf = it & 1;
      if (it >= 2) TC_TIMED(1, mbar_wait(acc_empty + buf, ((it >> 1) - 1) & 1));   // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)buf * (F16 ? 2u : 1u) * t.acc_stride;   // F16: D1 | D2 side by side
      uint32_t accumulate = 0;
      uint32_t blk = 0;                                            // running weight block index (resident image)
      const uint32_t nts_img = per_nt ? 1u : (uint32_t)t.n_tiles;   // N tiles interleaved in the shared-memory image
      const uint32_t nt_img = per_nt ? 0u : (uint32_t)w.nt;
      for (int ph = 0; ph < nph; ++ph) {
        const int K = ph == 0 ? a.ksize : 1, Cin = ph == 0 ? a.Cin : a.Cin2;
        const int CinP = (Cin + CB - 1) / CB * CB;
        const int HH = TC_TH + K - 1, HWP = K == 1 ? 8 : 16;
        const uint32_t a_lo_off = (uint32_t)HH * HWP * ROWB, b_lo_off = (uint32_t)t.n_tile * ROWB;
        for (int c0 = 0; c0 < CinP; c0 += CB) {
          const int as = a_it % A_ST;
          TC_TIMED(0, mbar_wait(a_full + as, (a_it / A_ST) & 1));
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + (size_t)as * t.a_stage_bytes);
          // One descriptor per operand per K-block; per tap only 16-B-unit offsets are added to the low word
          // (the issuing thread is alone on its scheduler: every dependent instruction costs ~5 cycles).
          const uint64_t ad0 = umma_desc(a_base, (uint32_t)HWP * ROWB, LAYOUT);
          const uint64_t bd_res0 = umma_desc(smem_u32(sB), 8 * ROWB, LAYOUT);
          const uint32_t a_lo16 = a_lo_off >> 4, b_lo16 = b_lo_off >> 4, blk16 = t.blk_bytes >> 4;
          uint32_t bidx = (blk + nt_img) * blk16;            // resident image: block of (K-block, tap 0, N tile)
          if (F16 && t.resident) {                                   // fast path: whole K-block from one elected region
            if (!b_ready) { TC_TIMED(2, mbar_wait(b_full, 0)); b_ready = true; tc_fence_after(); }
            const uint32_t d2 = d_tmem + (uint32_t)t.n_tile, bstride = nts_img * blk16;
            constexpr int KS = CB == 32 ? 2 : 1;
            if (elect_one()) {
              if (K == 3) umma_kblock_f16<ROWB, 3, KS>(d_tmem, d2, ad0, a_lo16, bd_res0 + bidx, bstride, idesc2, idesc, accumulate);
              else if (K == 7) umma_kblock_f16<ROWB, 7, KS>(d_tmem, d2, ad0, a_lo16, bd_res0 + bidx, bstride, idesc2, idesc, accumulate);
              else umma_kblock_f16<ROWB, 1, KS>(d_tmem, d2, ad0, a_lo16, bd_res0 + bidx, bstride, idesc2, idesc, accumulate);
            }
            accumulate = 1;
          } else
          for (int dy = 0; dy < K; ++dy) {
            int bs = 0;                                              // ring stage of this tap row (streamed weights)
            for (int dx = 0; dx < K; ++dx) {
              uint64_t bd_hi;
              if (!CAN_STREAM || t.resident) {
                if (!b_ready) { TC_TIMED(2, mbar_wait(b_full, 0)); b_ready = true; }
                bd_hi = bd_res0 + bidx;
                bidx += nts_img * blk16;
              } else {
                if (dx == 0) {
                  bs = b_it % B_ST;
                  TC_TIMED(2, mbar_wait(b_full + bs, (b_it / B_ST) & 1));
                }
                bd_hi = bd_res0 + (uint32_t)bs * (t.b_stage_bytes >> 4) + (uint32_t)dx * blk16;
              }
              tc_fence_after();
              const uint64_t ad_hi = ad0 + (uint32_t)((dy * HWP + dx) * (ROWB >> 4));
              const uint64_t ad_lo = ad_hi + a_lo16, bd_lo = bd_hi + b_lo16;
              if constexpr (F16) {                                   // K = 16 per instruction
                // A weight block is [B_hi rows ; B_lo rows] contiguously, so A_hi x [B_hi;B_lo] is ONE MMA of width
                // 2n writing D1 | D2 side by side; only A_lo x B_hi remains as a second (width n) MMA into D2:
                // two instructions per K step instead of three (the ~48-cycle per-instruction overhead dominates).
                const uint32_t d2 = d_tmem + (uint32_t)t.n_tile;
                (void)bd_lo;
                if (elect_one()) umma_tap_f16<CB == 32 ? 2 : 1>(d_tmem, d2, ad_hi, ad_lo, bd_hi, idesc2, idesc, accumulate);
              } else {                                               // 3xTF32, K = 8: lo*hi, hi*lo, hi*hi
                if (elect_one()) {
                  umma_tf32(d_tmem, ad_lo, bd_hi, idesc, accumulate);
                  umma_tf32(d_tmem, ad_lo + 2, bd_hi + 2, idesc, 1);
                  umma_tf32(d_tmem, ad_lo + 4, bd_hi + 4, idesc, 1);
                  umma_tf32(d_tmem, ad_lo + 6, bd_hi + 6, idesc, 1);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_tf32(d_tmem, ad_hi + 2 * ks, bd_lo + 2 * ks, idesc, 1);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_tf32(d_tmem, ad_hi + 2 * ks, bd_hi + 2 * ks, idesc, 1);
                }
              }
              accumulate = 1;
              if (CAN_STREAM && !t.resident && dx == K - 1) {
                if (elect_one()) umma_commit(b_empty + bs);          // the row's stage is reusable when its MMAs retire
                ++b_it;
              }
            }
          }
          blk += (uint32_t)K * K * nts_img;
          if (elect_one()) umma_commit(a_empty + as);
          ++a_it;
        }
      }
      if (elect_one()) umma_commit(acc_full + buf);
      ++it;
    }
    if (t.prof && (tid & 31) == 0) {
      atomicAdd(t.prof + 2, (unsigned long long)pc[0]);
      atomicAdd(t.prof + 3, (unsigned long long)pc[1]);
      atomicAdd(t.prof + 4, (unsigned long long)pc[2]);
      atomicAdd(t.prof + 5, (unsigned long long)(clock64() - mt0));
      atomicAdd(t.prof + 8, (unsigned long long)it);
    }
  } else if (warp >= LW + 2) {
    // =================================== epilogue (4 warps) ===========================================
    const int quarter = warp & 3;                                  // TMEM lanes this warp may read
    const int r = quarter * 32 + (tid & 31);                       // accumulator row = output pixel of the tile
    int it = 0;
    long long pc[4] = {0, 0, 0, 0};
    const long long et0 = clock64();
    for (int item = blockIdx.x; item < t.total_items; item += gridDim.x) {
      TcItem w;
      if (!tc_decode(t, item, w)) continue;
      const int buf = it & 1;
      TC_TIMED(0, mbar_wait(acc_full + buf, (it >> 1) & 1));
      tc_fence_after();
      const int oy = w.y0 + (r >> 3), ox = w.x0 + (r & 7);
      const bool px_ok = oy < a.H && ox < a.W;
      const int co_base = w.nt * t.n_tile;
      const uint32_t t_row = tmem + (uint32_t)buf * (F16 ? 2u : 1u) * t.acc_stride + ((uint32_t)(quarter * 32) << 16);
      // Output addressing, hoisted per item.  plain: NHWC pixel (oy, ox); upsample (ConvTranspose k2s2 as 1x1 to 4*Co +
      // pixel shuffle): column co = q*Co + c goes to pixel (2*oy + q/2, 2*ox + q%2), channel c - a 16-column chunk never
      // straddles q because Co is a multiple of 16; nchw: planar store of the real channels.
      const int Co = a.upsample ? (a.CoutP >> 2) : a.CoutP;
      const int Ho = a.upsample ? 2 * a.H : a.H, Wo = a.upsample ? 2 * a.W : a.W;
      const size_t img_px = (size_t)w.img * Ho * Wo;
      const size_t px_plain = img_px + (size_t)(a.upsample ? 2 * oy : oy) * Wo + (a.upsample ? 2 * ox : ox);
      auto chunk_offset = [&](int co, int& ch) -> size_t {        // float offset of column co's pixel (NHWC) ; ch = channel
        if (!a.upsample) { ch = co; return px_plain * a.CoutS; }
        const int q = co / Co;
        ch = co - q * Co;
        return (px_plain + (size_t)(q >> 1) * Wo + (q & 1)) * a.CoutS;
      };
      constexpr int CW = OCC == 2 ? 8 : 16, G4 = CW / 4;         // columns per chunk (8 keeps the 2-CTA variant in 72 registers)
      // Split tensors (fp16 hi plane, then the scaled-lo plane): residuals only reach the TMA variants, split outputs also
      // leave the 16-channel legacy variant (the 7x7 front layer feeds the first TMA layer).
      constexpr bool SPLIT_RES_OK = TMA, SPLIT_OUT_OK = TMA || MODE == 2;
      const bool res_split = SPLIT_RES_OK && (a.fmt & FVP_FMT_RES_SPLIT), out_split = SPLIT_OUT_OK && (a.fmt & FVP_FMT_OUT_SPLIT);
      const size_t out_plane = (size_t)a.n * Ho * Wo * a.CoutS;    // halves per plane of the output / residual tensor
      // residuals of one chunk, issued early and kept RAW (converted at their use): fp32 = G4 float4; split = CW/8 uint4 of
      // hi halves followed by CW/8 uint4 of lo halves
      auto fetch_res = [&](int cb, uint4* rr) {
#pragma unroll
        for (int g4 = 0; g4 < G4; ++g4) rr[g4] = make_uint4(0u, 0u, 0u, 0u);
        if (a.res_mode && px_ok && cb < t.n_tile) {
          int ch;
          const size_t off = chunk_offset(co_base + cb, ch);
          if (res_split) {
            const __half* rh = reinterpret_cast<const __half*>(a.res) + off + ch;
#pragma unroll
            for (int g8 = 0; g8 < CW / 8; ++g8)
              if (co_base + cb + g8 * 8 < a.CoutP) {
                rr[g8] = __ldg((const uint4*)(rh + g8 * 8));
                rr[CW / 8 + g8] = __ldg((const uint4*)(rh + out_plane + g8 * 8));
              }
          } else {
#pragma unroll
            for (int g4 = 0; g4 < G4; ++g4)
              if (co_base + cb + g4 * 4 < a.CoutP) rr[g4] = __ldg((const uint4*)(a.res + off + ch + g4 * 4));
          }
        }
      };
      auto res_value = [&](const uint4* rr, int i) -> float {      // residual of column i of the chunk (i is a constant after unrolling)
        if (res_split) {
          const uint32_t wh = reinterpret_cast<const uint32_t*>(&rr[i / 8])[(i % 8) / 2];
          const uint32_t wl = reinterpret_cast<const uint32_t*>(&rr[CW / 8 + i / 8])[(i % 8) / 2];
          const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&wh)), l = __half22float2(*reinterpret_cast<const __half2*>(&wl));
          return (i & 1) ? fmaf(l.y, 1.0f / 2048.0f, h.y) : fmaf(l.x, 1.0f / 2048.0f, h.x);
        }
        return __uint_as_float(reinterpret_cast<const uint32_t*>(&rr[i / 4])[i % 4]);
      };
      uint4 rnext[G4];
      fetch_res(0, rnext);
      float amax = 0.0f;                                           // range guard: largest stored magnitude of this item
      for (int cb = 0; cb < t.n_tile; cb += CW) {
        uint4 rcur[G4];
#pragma unroll
        for (int g4 = 0; g4 < G4; ++g4) rcur[g4] = rnext[g4];
        uint32_t u1[CW], u2[CW];
        const long long tl0 = t.prof ? clock64() : 0;
        tmem_ld_issue<CW>(t_row + (uint32_t)cb, u1);               // both accumulators in flight, one wait
        if constexpr (F16) tmem_ld_issue<CW>(t_row + (uint32_t)t.n_tile + (uint32_t)cb, u2);
        fetch_res(cb + CW, rnext);                                 // next chunk's residuals fly under this chunk
        tmem_ld_wait();
        if (t.prof) pc[1] += clock64() - tl0;
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          v[i] = __uint_as_float(u1[i]);
          if constexpr (F16) v[i] = fmaf(__uint_as_float(u2[i]), 1.0f / 2048.0f, v[i]);   // D = D1 + 2^-11 * D2
        }
        if (cb + CW >= t.n_tile) {                                 // last chunk read: hand the accumulator back
          tc_fence_before();
          mbar_arrive(acc_empty + buf);
        }
        if (!px_ok) continue;
        int ch0;
        const size_t off = chunk_offset(co_base + cb, ch0);
        // bias, residual, ReLU - in place in v[]
#pragma unroll
        for (int g4 = 0; g4 < G4; ++g4) {
          const int co = co_base + cb + g4 * 4;
          if (co >= a.CoutP) break;
          const float4 bias = *(const float4*)(s_bias + co);
          const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float o = v[g4 * 4 + e] + bb[e];
            if (a.res_mode == 1) o += res_value(rcur, g4 * 4 + e);
            if (a.relu) o = fmaxf(o, 0.f);
            if (a.res_mode == 2) o += res_value(rcur, g4 * 4 + e);
            v[g4 * 4 + e] = o;
          }
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[g4 * 4]), fabsf(v[g4 * 4 + 1])), fmaxf(fabsf(v[g4 * 4 + 2]), fabsf(v[g4 * 4 + 3]))));
        }
        if (out_split) {
          // one 16-byte store of 8 hi halves and one of 8 scaled-lo halves per 8 channels: the operand form the next
          // layer's TMA loads drop into shared memory as they are (x = hi + lo * 2^-11, 22 significant bits)
          __half* oh = reinterpret_cast<__half*>(a.out) + off + ch0;
#pragma unroll
          for (int g8 = 0; g8 < CW / 8; ++g8) {
            if (co_base + cb + g8 * 8 >= a.CoutP) break;
            uint32_t ph_[4], pl_[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = v[g8 * 8 + 2 * e], x1 = v[g8 * 8 + 2 * e + 1];
              const __half2 h2 = __floats2half2_rn(x0, x1);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn((x0 - hf.x) * 2048.0f, (x1 - hf.y) * 2048.0f);
              ph_[e] = *reinterpret_cast<const uint32_t*>(&h2);
              pl_[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            *(uint4*)(oh + g8 * 8) = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
            *(uint4*)(oh + out_plane + g8 * 8) = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
          }
          continue;
        }
#pragma unroll
        for (int g4 = 0; g4 < G4; ++g4) {
          const int co = co_base + cb + g4 * 4;
          if (co >= a.CoutP) break;
          const int ch = ch0 + g4 * 4;
          if (!a.nchw) {
            *(float4*)(a.out + off + ch) = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
          } else {                                                 // planar [n][CoutReal][Ho][Wo] (final layers, no upsample)
            const size_t plane = (size_t)Ho * Wo;
            float* op = a.out + (size_t)w.img * a.CoutReal * plane + (size_t)oy * Wo + ox;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (ch + e < a.CoutReal) op[(size_t)(ch + e) * plane] = v[g4 * 4 + e];
          }
        }
      }
      // Range guard (also in the 3xTF32 variant, whose outputs may feed an fp16 layer).  A maximum ignores NaN, but a NaN
      // needs an infinite operand, and the first value that cannot become a finite fp16 "hi" part is caught here.  (A test
      // per store instead of this per-item maximum measured 17 % slower on the 3x3 32->32 layer, profiles/r02_conv_ab.txt.)
      if (amax >= 65504.0f && g_tc_status) *g_tc_status = 1;
      ++it;
    }
    if (t.prof && tid == LW * 32 + 64) {
      atomicAdd(t.prof + 6, (unsigned long long)pc[0]);
      atomicAdd(t.prof + 7, (unsigned long long)(clock64() - et0));
      atomicAdd(t.prof + 9, (unsigned long long)pc[1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(t.tmem_cols));
  }
}

}  // namespace

// ---- TMA descriptors of split activation tensors -------------------------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver (libcuda); fetched at run time so that the library has no link-time
// dependency on libcuda (it must load - and export its symbols - on hosts without a driver).
typedef CUresult (*FvpEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static FvpEncodeTiled fvp_encode_tiled() {
  static const FvpEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (FvpEncodeTiled)p;
  }();
  return fn;
}
// one plane [n][H][W][C] of fp16; box = (cb channels, halo pitch, halo height, 1 image), swizzle = the K-block row size
static bool fvp_make_map(CUtensorMap* m, const void* base, int n, int H, int W, int C, int cb, int hwp, int hh) {
  FvpEncodeTiled enc = fvp_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)hwp, (cuuint32_t)hh, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// host: geometry of the tiled weight image (must match pack_tc in fvp_params.cu).  narrow = 32-column N tiles
// (more CTAs for small launches), otherwise up to 128 columns per CTA.  K-blocks are always 32 channels.
void fvp_tc_geometry(int coutp, int narrow, int* n_tile, int* n_tiles) {
  const int npad = fvp_round_up(coutp, 16);
  const int cap = narrow == 1 ? 32 : (narrow == 2 ? 64 : 128);     // 0: up to 128 columns, 1: 32, 2: 64
  *n_tile = npad <= cap ? npad : cap;
  *n_tiles = fvp_cdiv(npad, *n_tile);
}

// read-only after start-up: FVP_TC_OCC=1 never co-schedules two CTAs per SM (A/B switch of tools/gpu_conv_check.sh)
static const int g_tc_occ = [] { const char* e = std::getenv("FVP_TC_OCC"); return e ? std::atoi(e) : 2; }();

// Per-device setup (fvp_create): opt-in shared memory of all six instantiations, carve-out of the two-CTA variants.
cudaError_t fvp_conv_tc_init_device() {
  const void* fns[10] = {(const void*)k_conv_tc<0, 1, false>, (const void*)k_conv_tc<1, 1, false>, (const void*)k_conv_tc<2, 1, false>,
                         (const void*)k_conv_tc<0, 2, false>, (const void*)k_conv_tc<1, 2, false>, (const void*)k_conv_tc<2, 2, false>,
                         (const void*)k_conv_tc<1, 1, true>,  (const void*)k_conv_tc<2, 1, true>,
                         (const void*)k_conv_tc<1, 2, true>,  (const void*)k_conv_tc<2, 2, true>};
  const bool two_cta[10] = {false, false, false, true, true, true, false, false, true, true};
  for (int i = 0; i < 10; ++i) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, fns[i]);
    if (e != cudaSuccess) return e;
    if (fa.sharedSizeBytes > (size_t)TC_STATIC_SMEM) {
      fprintf(stderr, "fvp_conv_tc_init_device: static shared memory %zu B exceeds TC_STATIC_SMEM\n", fa.sharedSizeBytes);
      return cudaErrorInvalidValue;
    }
    e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_DYN_SMEM_MAX);
    if (e != cudaSuccess) return e;
    if (two_cta[i]) {
      e = cudaFuncSetAttribute(fns[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}
cudaError_t fvp_conv_tc_set_status_ptr(int* d_status) {
  return cudaMemcpyToSymbol(g_tc_status, &d_status, sizeof(d_status));
}

// mode: 0 = 3xTF32, 1 = fp16 split with 32-channel K-blocks, 2 = fp16 split with 16-channel K-blocks
// wtc[3]: weight images tiled for N tiles of up to 128 / 32 / 64 columns (NULL where not packed)
void fvp_launch_conv_tc(const FvpConvArgs& a, const float* const wtc[3], int mode, const FvpLaunchEnv& env, cudaStream_t st) {
  const int num_sms = env.num_sms;
  int* const plan_out = env.tc_plan;
  if ((mode != 2 && a.ksize > 3) || a.CoutP > TC_MAX_COUT) {     // the loaders of the 32-channel K-block variants hold a 3x3 halo at most (EPT); 7x7 layers
    if (plan_out) { plan_out[0] = 0; return; }        // plan query: "not handled by the tensor-core engine"
    fvp_launch_conv(a, st);           // of this network run in mode 2 (16-channel K-blocks)
    return;
  }
  TcArgs t;
  t.c = a;
  t.prof = env.tc_prof;
  const uint32_t rowb = mode == 0 ? 128 : (mode == 1 ? 64 : 32);
  const int cb = mode == 2 ? 16 : 32, f16 = mode != 0;
  const int tiles = fvp_cdiv(a.H, TC_TH) * fvp_cdiv(a.W, TC_TW) * a.n;
  // N-tile width: the persistent grid runs ceil(items / SMs) rounds of items; per K step an item costs roughly
  // (48 + fetch + math) cycles per MMA.  Pick the width that minimises rounds x per-item cost.
  const int k = a.ksize;
  const uint32_t a0 = (uint32_t)(TC_TH + k - 1) * (k == 1 ? 8 : 16) * rowb * 2;
  const uint32_t a1 = a.in2 ? (uint32_t)TC_TH * 8 * rowb * 2 : 0;
  t.a_stage_bytes = a0 > a1 ? a0 : a1;                            // multiples of 1024
  const int nblocks = k * k * (fvp_round_up(a.Cin, cb) / cb) + (a.in2 ? fvp_round_up(a.Cin2, cb) / cb : 0);
  const uint32_t budget = TC_DYN_SMEM_MAX - 1024;                  // dynamic smem we may use (1 KB alignment slack)
  // Weight residency of a variant: 1 = the whole image fits beside a double-buffered halo; 2 = only one N tile's blocks fit
  // (small launches: every CTA then serves a single N tile); 0 = streamed per tap row (charged 1.4x the per-MMA cost of the
  // unrolled resident issue path; 1.8x before the row-wise ring - with 1.4 the 128-channel layers of P2PNet at batch 1 take
  // 64-column streamed tiles instead of 32-column resident ones: measured 0.351 -> 0.338 ms for the trunk,
  // profiles/r02_bench_b1_serial_ntile64.json, while CenterNet's single image keeps the 32-column tiles).
  auto residency = [&](int nt, int nts) -> int {
    const uint32_t blk = (uint32_t)nt * rowb * 2;
    if ((uint64_t)nblocks * nts * blk + 2 * t.a_stage_bytes <= budget) return 1;
    if (f16 && nts > 1 && num_sms >= nts && (uint64_t)nblocks * blk + 2 * t.a_stage_bytes <= budget &&
        fvp_cdiv(tiles * nts, num_sms - num_sms % nts) <= 2)
      return 2;
    return 0;
  };
  int best = 0;
  double best_cost = 1e30;
  for (int v = 0; v < 3; ++v) {
    if (!wtc[v]) continue;
    int nt, nts;
    fvp_tc_geometry(a.CoutP, v, &nt, &nts);
    const double per_item = f16 ? (48 + (128 + 2 * nt) / 4.0 + nt) + (48 + (128 + nt) / 4.0 + nt / 2.0)
                                : 3.0 * (48 + (128 + nt) / 4.0 + nt / 2.0);
    const double cost = (double)fvp_cdiv(tiles * nts, num_sms) * per_item * (residency(nt, nts) ? 1.0 : 1.4) + 1.0 * nts;   // ties: wider tile
    if (cost < best_cost) { best_cost = cost; best = v; }
  }
  fvp_tc_geometry(a.CoutP, best, &t.n_tile, &t.n_tiles);
  t.wtc = wtc[best];
  t.blk_bytes = (uint32_t)t.n_tile * rowb * 2;
  const uint32_t image_bytes = (uint32_t)nblocks * t.n_tiles * t.blk_bytes;
  const int res = residency(t.n_tile, t.n_tiles);
  const uint32_t stage_bytes = (uint32_t)k * t.blk_bytes;         // streamed weights: one ring stage per tap row
  const int min_stages = 2;
  const int stream2 = (2 * t.a_stage_bytes + min_stages * stage_bytes <= budget) ? (int)((budget - 2 * t.a_stage_bytes) / stage_bytes) : 0;
  if (res == 1) {                                                  // weights resident, double-buffered halo
    t.resident = 1; t.a_stages = 2; t.b_stages = 1; t.b_stage_bytes = image_bytes;
  } else if (res == 2) {                                           // one N tile's weights resident per CTA
    t.resident = 2; t.a_stages = 2; t.b_stages = 1; t.b_stage_bytes = (uint32_t)nblocks * t.blk_bytes;
  } else if (stream2 >= min_stages) {                              // streamed weights, double-buffered halo
    t.resident = 0; t.a_stages = 2; t.b_stage_bytes = stage_bytes;
    t.b_stages = stream2 > TC_MAX_B ? TC_MAX_B : stream2;
  } else if (image_bytes + t.a_stage_bytes <= budget) {
    t.resident = 1; t.a_stages = 1; t.b_stages = 1; t.b_stage_bytes = image_bytes;
  } else {
    t.resident = 0; t.a_stages = 1; t.b_stage_bytes = stage_bytes;
    const int bs = (int)((budget - t.a_stage_bytes) / stage_bytes);
    t.b_stages = bs > TC_MAX_B ? TC_MAX_B : bs;
  }
  t.acc_stride = (uint32_t)fvp_round_up(t.n_tile, 32);          // accumulators side by side in TMEM: 2 buffers x (D1[,D2])
  t.tmem_cols = 32;
  while (t.tmem_cols < (f16 ? 4u : 2u) * t.acc_stride) t.tmem_cols <<= 1;
  t.tiles_x = fvp_cdiv(a.W, TC_TW);
  t.tiles_per_img = t.tiles_x * fvp_cdiv(a.H, TC_TH);
  t.total_items = t.tiles_per_img * a.n * t.n_tiles;
  auto magic = [&](int d) -> uint32_t {            // d == 1 would need 2^32: the plain division handles it
    return (d > 1 && (unsigned long long)t.total_items * d < (1ull << 32)) ? (uint32_t)(((1ull << 32) + d - 1) / d) : 0u;
  };
  t.m_ntiles = magic(t.n_tiles); t.m_tpi = magic(t.tiles_per_img); t.m_tx = magic(t.tiles_x);
  const size_t smem = 1024 + (size_t)t.a_stages * t.a_stage_bytes + (size_t)t.b_stages * t.b_stage_bytes;
  // Two CTAs per SM when both fit: 228 KB of shared memory per SM, 1 KB reserved per CTA, TC_STATIC_SMEM static; 512 TMEM columns.
  // (A CTA that needs more than 256 columns must never share an SM with a sibling: its tcgen05.alloc would block.)
  // Measured per layer (profiles/r01_s5_conv_layers_occ*.txt, 960 images): 7x7 16->16 1.49x, 1x1 32->16 1.43x, 3x3 32->32 1.08x
  // faster with two CTAs per SM; the 3x3 16->32 layer (16-channel K-blocks, loader-bound) is 9 % slower and keeps one.
  // (Also tried and dropped: one LDG.128 per lane over whole 128-B lines with 8-byte shared stores - fewer L1 sector
  //  lookups but twice the store instructions: 3 % slower over the trunk's layer mix.)
  // split (TMA-fed) input: no loader warps, so the loader-bound exception of the 16-channel 3x3 layer does not apply
  const bool tma = (a.fmt & FVP_FMT_IN_SPLIT) != 0;
  const bool occ2 = t.resident && g_tc_occ != 1 && !(!tma && mode == 2 && k == 3 && g_tc_occ != 3) && 2 * (smem + TC_STATIC_SMEM + 1024) <= 228 * 1024 && t.tmem_cols <= 256 && t.total_items > num_sms;
  const int slots = num_sms * (occ2 ? 2 : 1);
  int grid = t.total_items < slots ? t.total_items : slots;              // persistent: one or two CTAs per SM
  if (t.resident == 2) grid -= grid % t.n_tiles;                         // CTA b serves N tile b % n_tiles only
  if (plan_out) {                                                        // host-only query of the decisions above (tests)
    const int plan[10] = {t.n_tile, t.n_tiles, t.resident, t.a_stages, t.b_stages, (int)smem, occ2 ? 1 : 0, grid, t.total_items,
                          (int)t.tmem_cols};
    for (int i = 0; i < 10; ++i) plan_out[i] = plan[i];
    return;
  }
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (tma) {
    if (mode == 0) {                   // split tensors are the operand format of the fp16 engine: a caller bug, reported, never launched
      fprintf(stderr, "fvp_launch_conv_tc: split input needs the fp16 engine\n");
      if (env.error) *env.error = 1;
      return;
    }
    const int hh = TC_TH + k - 1, hwp = k == 1 ? 8 : 16;
    const __half* in_hi = reinterpret_cast<const __half*>(a.in);
    bool ok = fvp_make_map(&maps.in_hi, in_hi, a.n, a.H, a.W, a.Cin, cb, hwp, hh) &&
              fvp_make_map(&maps.in_lo, in_hi + (size_t)a.n * a.H * a.W * a.Cin, a.n, a.H, a.W, a.Cin, cb, hwp, hh);
    if (ok && a.in2) {
      const __half* in2_hi = reinterpret_cast<const __half*>(a.in2);
      ok = fvp_make_map(&maps.in2_hi, in2_hi, a.n, a.H, a.W, a.Cin2, cb, 8, TC_TH) &&
           fvp_make_map(&maps.in2_lo, in2_hi + (size_t)a.n * a.H * a.W * a.Cin2, a.n, a.H, a.W, a.Cin2, cb, 8, TC_TH);
    }
    if (!ok) {                         // no descriptor, no launch: the entry point that built `env` turns the mark into FVP_E_CUDA
      fprintf(stderr, "fvp_launch_conv_tc: cuTensorMapEncodeTiled failed (n=%d %dx%d C=%d/%d k=%d)\n", a.n, a.H, a.W, a.Cin, a.Cin2, k);
      if (env.error) *env.error = 1;
      return;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tma ? 192 : TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (tma) {
    if (occ2) {
      if (mode == 2) cudaLaunchKernelEx(&cfg, k_conv_tc<2, 2, true>, t, maps);
      else cudaLaunchKernelEx(&cfg, k_conv_tc<1, 2, true>, t, maps);
    } else {
      if (mode == 2) cudaLaunchKernelEx(&cfg, k_conv_tc<2, 1, true>, t, maps);
      else cudaLaunchKernelEx(&cfg, k_conv_tc<1, 1, true>, t, maps);
    }
  } else if (occ2) {
    if (mode == 2) cudaLaunchKernelEx(&cfg, k_conv_tc<2, 2, false>, t, maps);
    else if (mode == 1) cudaLaunchKernelEx(&cfg, k_conv_tc<1, 2, false>, t, maps);
    else cudaLaunchKernelEx(&cfg, k_conv_tc<0, 2, false>, t, maps);
  } else {
    if (mode == 2) cudaLaunchKernelEx(&cfg, k_conv_tc<2, 1, false>, t, maps);
    else if (mode == 1) cudaLaunchKernelEx(&cfg, k_conv_tc<1, 1, false>, t, maps);
    else cudaLaunchKernelEx(&cfg, k_conv_tc<0, 1, false>, t, maps);
  }
}

// Host-only: the launch plan fvp_launch_conv_tc would choose for one layer (no CUDA call is made).  out[10] = n_tile, n_tiles,
// weight residency (0 streamed, 1 whole image, 2 one N tile), A stages, B stages, dynamic smem bytes, two CTAs per SM (0/1),
// grid, work items, TMEM columns; out[0] = 0 when the layer falls back to the CUDA-core kernel.  engine: 2 = fp16 split
// (16-channel K-blocks when cin <= 16 like the trunk dispatch in fvp_conv.cu), 1 = 3xTF32.
extern "C" int fvp_debug_conv_plan(int n, int H, int W, int cin, int cin2, int cout, int k, int engine, int num_sms, int* out) {
  if (!out || n <= 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0 || (k != 1 && k != 3 && k != 7) || num_sms <= 0) return -1;
  FvpConvArgs a = {};
  static const float dummy = 0.f;                    // only the NULL-ness of the pointers matters to the planner
  a.n = n; a.H = H; a.W = W; a.Cin = cin; a.Cin2 = cin2; a.in2 = cin2 ? &dummy : nullptr;
  a.CoutP = fvp_round_up(cout, 4); a.ksize = k;
  const int npad = fvp_round_up(a.CoutP, 16);
  const float* wide[3] = {&dummy, npad > 32 ? &dummy : nullptr, npad > 64 ? &dummy : nullptr};     // as stash() packs them
  const float* c16[3] = {&dummy, nullptr, nullptr};
  const bool use_c16 = engine == 2 && ((cin <= 16 && cin2 <= 16) || (k == 7 && cin <= 32));     // as stash() packs them
  for (int i = 0; i < 10; ++i) out[i] = 0;
  const FvpLaunchEnv env{num_sms, engine, nullptr, out, 0, nullptr};
  fvp_launch_conv_tc(a, use_c16 ? c16 : wide, engine == 2 ? (use_c16 ? 2 : 1) : 0, env, nullptr);
  return 0;
}
