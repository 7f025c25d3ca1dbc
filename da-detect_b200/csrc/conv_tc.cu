// tcgen05 / TMA arm of the dense tier (impl = DD_IMPL_TCGEN05), sm_100a only.
//
// Implicit-GEMM convolution on the 5th-generation tensor cores: fp32 NHWC activations and OHWI weights
// are consumed directly as TF32 operands (kind::tf32, fp32 accumulate in TMEM) — no im2col buffer, no
// layout or precision conversion pass.  One warp-specialised CTA computes a 128 x BN output tile:
//
//   warp 0   TMA producer : per (filter tap, 32-channel slice) one 4-D tensor-map box of the activation
//                           (a [tn x th x tw] block of output pixels shifted by the tap; out-of-image
//                           coordinates are zero-filled by TMA = the conv padding) and one 2-D box of
//                           the weights, both 128B-swizzled K-major tiles, into a 3-4 stage smem ring
//   warp 1   MMA issuer   : one elected thread issues 4 tcgen05.mma (M=128, N=BN, K=8) per stage,
//                           tcgen05.commit releases the stage back to the producer
//   warps 2-5 epilogue    : tcgen05.ld the fp32 accumulators (one output pixel per thread, 32 channels per
//                           load), apply FrozenBN scale/bias + residual + ReLU (forward) or fan-in add +
//                           ReLU mask (dgrad), 128-bit stores to NHWC
//
//   forward: D[pix, co] = sum_{tap, ci} X[pix + tap, ci] * W[co, tap, ci]
//   dgrad  : the same kernel on GY with the flipped / transposed / BN-scaled weights W'[ci, tap', co]
//            (prepared by a small transpose kernel); 1x1 stride-2 dgrad runs compact and scatters rows
//   wgrad  : D[co, ci] (per tap) = sum_pix GY[pix, co] * X[pix + tap, ci]: both operands MN-major
//            (channels contiguous), split over pixel blocks, partials reduced by wgrad_reduce
//
// Reference being replaced: cuDNN/cuBLAS calls behind F.conv2d / F.linear.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int BM = 128;            // output pixels per CTA tile (UMMA M)
constexpr int BKB = 128;           // bytes of K per stage row (one 128B swizzle span) = 32 tf32
constexpr int BKE = 32;            // elements of K per stage
constexpr int UMMA_K = 8;          // tf32: 32 bytes per MMA K step
constexpr int NUM_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int NUM_THREADS_X3 = 320;   // + warps 6..9: operand splitter of the 3xTF32 mode

struct TcParams {
  // epilogue
  float* out;
  const float* scale;     // per output channel (fwd: BN scale; may be null)
  const float* bias;      // per output channel (may be null)
  const float* extra;     // residual (fwd) / addend (dgrad), same addressing as out (may be null)
  const float* mask;      // dgrad: activation whose > 0 gates the result (may be null)
  int relu;
  // geometry of the GEMM rows (output pixel blocks)
  int N, OH, OW;          // logical output pixel grid enumerated by the tiles
  int tn, th, tw;         // tile = tn images x th rows x tw cols (tn*th*tw == 128)
  int tiles_h, tiles_w;   // ceil(OH/th), ceil(OW/tw)
  int m_tiles, n_tiles;   // persistent tile space: t -> (n_tile = t % n_tiles, m_tile = t / n_tiles)
  int tma_store;          // 1: staged chunks leave through a TMA tensor store; 0: guarded scalar stores
  int b_lo_row;           // X3: row offset of the low-part plane inside the prepared B matrix
  int stem;               // 1: A is the 5-D overlapping-window view of the padded NHWC4 image (7x7/2 stem)
  int prefetch_side;      // 1: map_e / map_m are valid and the producer prefetches extra / mask tiles to L2
  // where a row lands in the output tensor: out[((n*out_H + h*os)*out_W + w*os)*ldc + co]
  int out_H, out_W, os, ldc;
  int Cout;               // valid output channels (columns)
  // K loop
  int taps, KW, pad, cblocks;   // taps = KH*KW, cblocks = Cin/32
  int Cin;
  // 3xTF32, small maps (7x7 ROI maps fill 98 of the 128 rows of a {2, 8, 8} tile): the pixels of all maps are ONE dense
  // axis like for a 1x1 conv, a filter tap is a constant row offset (kh - pad) * flat_w + (kw - pad) of the TMA load,
  // and the operand splitter zeroes the rows whose tap falls outside their own flat_h x flat_w map (0 = off)
  int flat_h, flat_w;
  int onepass;             // 3xTF32 + residual + short K: one-pass epilogue (map_e = the residual's per-warp sub-box)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// ---- CTA-pair (cta_group::2) forms.  Both CTAs of a pair issue their own TMA loads; the transaction bytes are
// counted on the LEADER's (even CTA's) barrier: clearing bit 24 of a shared-window address selects the even CTA of
// the pair (CUTLASS Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// multicast forms (cluster of single-CTA MMAs sharing a B tile): the box lands at the same smem offset in every CTA
// of the mask and completes bytes on the barrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait with acquire at cluster scope: the arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the leader's MMAs: arrives on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TENSOR MEMORY (M = 128 rows = the 128 lanes, one 32-bit K element per column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-converged issue forms.  The MMA warp runs its loop with all 32 lanes converged (descriptor arithmetic is
// warp-uniform, so the compiler keeps it on the uniform datapath) and ONE elected lane issues each tcgen05
// instruction: `if (lane == 0) { ... }` around plain asm made the compiler wrap every operand of every MMA in an
// ELECT + R2UR sequence — ~25 SASS instructions per MMA, 300 per K iteration of the 3xTF32 kernel, which made the
// single issuing warp (not the tensor pipe, not shared memory, not L2) the limiter at ~1470 cycles per K iteration.
// Shared-memory descriptors are passed as their low word (start address >> 4, plus LBO bit 16 where needed); the
// high word is the compile-time constant of the layout.
constexpr uint32_t kDescHiKmajorSw128 = (uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29));   // SBO | version | swizzle
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_tf32_ss_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 ad, bd;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 ad, {%1, %5};\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32_ss_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 ad, bd;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 ad, {%1, %5};\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::tf32 [%0], ad, bd, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 bd;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One K iteration (32 channels = 4 K steps) of the 3xTF32 kernel as ONE asm block: 12 MMAs with the A operand in
// tensor memory (a_hi: 32 columns, a_lo = a_hi + 32) against the B_hi / B_lo tiles in shared memory, then the two
// commits (smem stage free; TMEM A slot free).  A single block lets ptxas keep the descriptor arithmetic on the
// uniform datapath and elect once.  MCAST: the stage commit is multicast to both CTAs of a multicast pair.
#define DD_X3_KITER_HEAD                                                                                            \
  "{\n\t"                                                                                                           \
  ".reg .pred e, p, t;\n\t"                                                                                         \
  ".reg .b32 ah1, ah2, ah3, al0, al1, al2, al3, x;\n\t"                                                             \
  ".reg .b64 bh0, bh1, bh2, bh3, bl0, bl1, bl2, bl3;\n\t"                                                           \
  "elect.sync _|e, 0xffffffff;\n\t"                                                                                 \
  "setp.ne.b32 p, %5, 0;\n\t"                                                                                       \
  "setp.eq.b32 t, 0, 0;\n\t"                                                                                        \
  "add.u32 ah1, %1, 8;\n\tadd.u32 ah2, %1, 16;\n\tadd.u32 ah3, %1, 24;\n\t"                                         \
  "add.u32 al0, %1, 32;\n\tadd.u32 al1, %1, 40;\n\tadd.u32 al2, %1, 48;\n\tadd.u32 al3, %1, 56;\n\t"                \
  "mov.b64 bh0, {%2, %6};\n\tadd.u32 x, %2, 2;\n\tmov.b64 bh1, {x, %6};\n\t"                                        \
  "add.u32 x, %2, 4;\n\tmov.b64 bh2, {x, %6};\n\tadd.u32 x, %2, 6;\n\tmov.b64 bh3, {x, %6};\n\t"                    \
  "mov.b64 bl0, {%3, %6};\n\tadd.u32 x, %3, 2;\n\tmov.b64 bl1, {x, %6};\n\t"                                        \
  "add.u32 x, %3, 4;\n\tmov.b64 bl2, {x, %6};\n\tadd.u32 x, %3, 6;\n\tmov.b64 bl3, {x, %6};\n\t"                    \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [al0], bh0, %4, p;\n\t" /* small terms first */                     \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bl0, %4, t;\n\t"                                              \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bh0, %4, t;\n\t"                                              \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [al1], bh1, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], bl1, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], bh1, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [al2], bh2, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], bl2, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], bh2, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [al3], bh3, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], bl3, %4, t;\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], bh3, %4, t;\n\t"
#define DD_X3_KITER_TAIL                                                                                            \
  "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"                              \
  "}"
template <bool MCAST>
__device__ __forceinline__ void umma_x3_kiter(uint32_t d, uint32_t a_hi, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate, uint64_t* stage_bar, uint64_t* aslot_bar) {
  if constexpr (MCAST) {
    asm volatile(DD_X3_KITER_HEAD
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %9;\n\t"
                 DD_X3_KITER_TAIL
                 ::"r"(d), "r"(a_hi), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128),
                   "r"(smem_u32(stage_bar)), "r"(smem_u32(aslot_bar)), "h"((uint16_t)3)
                 : "memory");
  } else {
    asm volatile(DD_X3_KITER_HEAD
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
                 DD_X3_KITER_TAIL
                 ::"r"(d), "r"(a_hi), "r"(b_hi), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128),
                   "r"(smem_u32(stage_bar)), "r"(smem_u32(aslot_bar))
                 : "memory");
  }
}

// The same for the plain TF32 kernel: 4 MMAs (A and B tiles in shared memory) and the stage commit.
// MODE 0: single CTA; 1: CTA pair (cta_group::2, commit multicast to both); 2: multicast pair (cta_group::1 MMAs,
// commit multicast to both).
#define DD_TF32_KITER_BODY(CG)                                                                                      \
  "{\n\t"                                                                                                           \
  ".reg .pred e, p, t;\n\t"                                                                                         \
  ".reg .b32 x;\n\t"                                                                                                \
  ".reg .b64 a0, a1, a2, a3, b0, b1, b2, b3;\n\t"                                                                   \
  "elect.sync _|e, 0xffffffff;\n\t"                                                                                 \
  "setp.ne.b32 p, %4, 0;\n\t"                                                                                       \
  "setp.eq.b32 t, 0, 0;\n\t"                                                                                        \
  "mov.b64 a0, {%1, %5};\n\tadd.u32 x, %1, 2;\n\tmov.b64 a1, {x, %5};\n\t"                                          \
  "add.u32 x, %1, 4;\n\tmov.b64 a2, {x, %5};\n\tadd.u32 x, %1, 6;\n\tmov.b64 a3, {x, %5};\n\t"                      \
  "mov.b64 b0, {%2, %5};\n\tadd.u32 x, %2, 2;\n\tmov.b64 b1, {x, %5};\n\t"                                          \
  "add.u32 x, %2, 4;\n\tmov.b64 b2, {x, %5};\n\tadd.u32 x, %2, 6;\n\tmov.b64 b3, {x, %5};\n\t"                      \
  "@e tcgen05.mma.cta_group::" CG ".kind::tf32 [%0], a0, b0, %3, p;\n\t"                                            \
  "@e tcgen05.mma.cta_group::" CG ".kind::tf32 [%0], a1, b1, %3, t;\n\t"                                            \
  "@e tcgen05.mma.cta_group::" CG ".kind::tf32 [%0], a2, b2, %3, t;\n\t"                                            \
  "@e tcgen05.mma.cta_group::" CG ".kind::tf32 [%0], a3, b3, %3, t;\n\t"
template <int MODE>
__device__ __forceinline__ void umma_tf32_kiter(uint32_t d, uint32_t a_d, uint32_t b_d, uint32_t idesc,
                                                uint32_t accumulate, uint64_t* stage_bar) {
  if constexpr (MODE == 1) {
    asm volatile(DD_TF32_KITER_BODY("2")
                 "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n\t}"
                 ::"r"(d), "r"(a_d), "r"(b_d), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128),
                   "r"(smem_u32(stage_bar)), "h"((uint16_t)3)
                 : "memory");
  } else if constexpr (MODE == 2) {
    asm volatile(DD_TF32_KITER_BODY("1")
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n\t}"
                 ::"r"(d), "r"(a_d), "r"(b_d), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128),
                   "r"(smem_u32(stage_bar)), "h"((uint16_t)3)
                 : "memory");
  } else {
    asm volatile(DD_TF32_KITER_BODY("1")
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
                 ::"r"(d), "r"(a_d), "r"(b_d), "r"(idesc), "r"(accumulate), "r"(kDescHiKmajorSw128),
                   "r"(smem_u32(stage_bar))
                 : "memory");
  }
}

__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit2_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc_e(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, 128B swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
// Rows are 128 B apart, 8-row swizzle atoms are 1024 B apart (SBO); LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// X3 = 3xTF32 ("fp32-grade") mode: every operand is split into a TF32-exact high part and a TF32-exact low part,
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi; a stage then holds four tiles [A_hi | A_lo | B_hi | B_lo].
// CTA2 = CTA-pair mode (tcgen05 cta_group::2): two CTAs of a cluster compute a 256 x BN tile; each holds its own 128
// rows of A and HALF of the B tile (BN/2 columns), the tensor cores of both SMs read both halves.
// PAIR: 0 = single CTAs; 1 = CTA pair with cta_group::2 MMAs (above); 2 = cluster of two single-CTA-MMA CTAs that
// share the B tile: each loads half of it and TMA-multicasts the half into both (half the B traffic from L2, the
// TMA -> splitter -> MMA chain stays inside the CTA).
template <int BN, bool X3, int PAIR, bool OP = false>
struct SmemLayout {
  static constexpr bool CTA2 = PAIR == 1;
  static constexpr int kABytes = BM * BKB;                       // 16 KB
  static constexpr int kBBytes = (CTA2 ? BN / 2 : BN) * BKB;     // 4 .. 32 KB
  // 3xTF32: the stage holds the RAW fp32 A tile and the pre-split B_hi / B_lo tiles; the split A operand (hi and lo
  // parts) lives in tensor memory, written there by the splitter warps and read from there by the MMAs
  static constexpr int kStageBytes = X3 ? kABytes + 2 * kBBytes : kABytes + kBBytes;
  static_assert(!(X3 && PAIR == 1), "the 3xTF32 kernel does not pair its MMAs (cta_group::2)");
  // one-pass epilogue (OP): per epilogue warp two 4 KB residual slices landed by TMA and a 1 KB scale / bias table
  static constexpr int kResBytes = OP ? 4 * 2 * 4096 : 0;
  static constexpr int kTabBytes = OP ? 4 * 1024 : 0;
  static constexpr int kFit = (192 * 1024 - kResBytes - kTabBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kStagingBytes = BM * 128;                 // one 128-row x 32-column fp32 chunk
  static constexpr int kStagingOff = kStages * kStageBytes;      // 2 staging buffers (1024-aligned)
  static constexpr int kResOff = kStagingOff + 2 * kStagingBytes;   // (1024-aligned: the slices are 128B-swizzled)
  static constexpr int kTabOff = kResOff + kResBytes;
  static constexpr int kRowOff = kTabOff + kTabBytes;            // int32 pixel index per tile row
  static constexpr int kBarOff = kRowOff + BM * 4;
  static constexpr int kTotal = kBarOff + 512 /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(kTotal <= 232448, "shared memory budget");
  static_assert(kStages >= 2, "at least two pipeline stages");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void load_scale_bias(const TcParams& p, int col, float (&sc)[4], float (&bi)[4]) {
  if (col + 3 < p.Cout && (p.Cout & 3) == 0) {           // aligned 128-bit loads (all multi-of-4 channel counts)
    const float4 s4 = p.scale ? dd::ldg4(p.scale + col) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b4 = p.bias ? dd::ldg4(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
    bi[0] = b4.x; bi[1] = b4.y; bi[2] = b4.z; bi[3] = b4.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool cok = col + e < p.Cout;
      sc[e] = (p.scale && cok) ? __ldg(p.scale + col + e) : 1.f;
      bi[e] = (p.bias && cok) ? __ldg(p.bias + col + e) : 0.f;
    }
  }
}

enum { EPI_EXTRA = 1, EPI_MASK = 2, EPI_SCALAR = 4, EPI_ONEPASS = 8 };

// ------------------------------------------------------------------------------------------------ kernel
// Persistent: grid = min(#tiles, #SMs); every role walks the same static tile sequence t = blockIdx.x + i*gridDim.x.
// Two TMEM accumulators (2 x BN columns) let the MMAs of tile i+1 run while the epilogue drains tile i; the TMA
// producer runs ahead across tile boundaries through the smem ring.
//
// Epilogue per 32-column chunk: tcgen05.ld (thread = one tile row) -> 128B-swizzled smem staging ->
// re-mapped pass (8 threads per row => coalesced 128-bit reads of residual / mask, per-channel scale + bias,
// ReLU) -> one TMA tensor store of the [rows x 32 ch] box, which also clips the tile against the tensor edges.
//
// CTA2 (CTA pair, cluster of 2): the pair walks tiles of 256 pixels x BN columns; CTA r owns pixel tile 2*mp + r (its
// own A loads, its own 128 TMEM lanes, its own epilogue) and loads columns [r*BN/2, (r+1)*BN/2) of the B tile.  Only
// the leader (rank 0) issues MMAs (cta_group::2, M = 256): half the B traffic per output and half the B operand
// reads per SM.  Barrier protocol of the pair:
//   full      leader's barrier counts the TMA bytes of BOTH CTAs (3xTF32: only the B halves; the A tile of each CTA
//             lands on that CTA's own barrier, where its splitter warps wait)
//   split     (3xTF32) leader's barrier, 8 arrivals: the splitter warps of both CTAs (remote arrive for rank 1)
//   empty     per CTA; the leader's tcgen05.commit is multicast to both
//   tmem_full per CTA, multicast commit; tmem_empty leader's barrier, 8 arrivals (epilogue warps of both CTAs)
template <int BN, int EPI, bool X3, int PAIR>
__global__ void __launch_bounds__(X3 ? NUM_THREADS_X3 : NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_e,
               const __grid_constant__ CUtensorMap map_m, const TcParams p) {
  constexpr bool kOnePass = (EPI & EPI_ONEPASS) != 0;
  static_assert(!kOnePass || (X3 && (EPI & (EPI_MASK | EPI_SCALAR)) == 0), "one-pass epilogue: 3xTF32, no mask input");
  using L = SmemLayout<BN, X3, PAIR, kOnePass>;
  constexpr bool CTA2 = PAIR == 1;     // cta_group::2 MMAs issued by the leader
  constexpr bool MC = PAIR == 2;       // own MMAs, B halves multicast between the two CTAs
  constexpr bool CL = PAIR != 0;       // launched as a cluster of 2
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + L::kStagingOff;
  int* row_pix = reinterpret_cast<int*>(smem + L::kRowOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full_bar = empty_bar + L::kStages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint64_t* split_bar = tmem_empty_bar + 2;             // [2], X3 only: the A operand of a TMEM slot has been written
  uint64_t* aslot_bar = split_bar + 2;                  // [2], X3 only: the MMAs that read a TMEM A slot have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aslot_bar + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff + 384);   // one-pass epilogue: [warp][buffer]
  constexpr int kBOff = L::kABytes;                     // B (hi) tile offset inside a stage; B_lo follows it
  constexpr int kBRows = CTA2 ? BN / 2 : BN;            // B rows (output columns) this CTA loads

  // the warp index through a shuffle: the compiler then KNOWS it is warp-uniform, treats the role branches below as
  // convergent and keeps the MMA warp's descriptor arithmetic in uniform registers (no ELECT / R2UR per operand)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int k_iters = p.taps * p.cblocks;
  const uint32_t rank = CL ? cluster_ctarank() : 0u;
  const bool leader = rank == 0 || MC;                 // MC: every CTA issues its own MMAs and owns its barriers
  // tile walk: a unit is one CTA or one pair
  const int unit = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = CL ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  const int total_tiles = m_units * p.n_tiles;
  // this CTA's 128-pixel tile of unit tile t: (n_tile, mt); mt >= m_tiles (odd count, rank 1) is a phantom tile whose
  // loads are zero-filled and whose stores are clipped away by TMA
  auto tile_of = [&](int t, int& n_tile, int& n0, int& oh0, int& ow0) {
    n_tile = t % p.n_tiles;
    int mt = t / p.n_tiles;
    if (CL) mt = 2 * mt + (int)rank;
    const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
    const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
    n0 = mt * p.tn; oh0 = tile_h * p.th; ow0 = tile_w * p.tw;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
    // MC: a stage is free when the MMAs of BOTH CTAs that read it have retired (the peer multicasts into it)
    for (int s = 0; s < L::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, MC ? 2 : 1); }
    // consumer-side barriers count WARPS (one elected arrive per warp after __syncwarp): a remote arrive is a DSMEM
    // transaction, and 128 of them per K-iteration made the pair mode a third slower than single CTAs
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar + s, 1); mbar_init(tmem_empty_bar + s, CTA2 ? 8 : 4); }
    if (X3) for (int s = 0; s < 2; ++s) { mbar_init(split_bar + s, 4); mbar_init(aslot_bar + s, 1); }
    if (kOnePass) for (int s = 0; s < 8; ++s) mbar_init(res_bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM: 2 x BN columns (two tile accumulators, ping-pong).  3xTF32 mode: [M | P0 | P1], BN columns each.
  // The tensor core's fp32 accumulator rounds toward zero on every accumulation step, a bias that grows with the
  // length of the accumulation chain (measured ~1e-5 relative after 432 steps).  So the MMA warp accumulates only
  // kGroup K-iterations (96 steps) into a partial slot P0/P1 (ping-pong), and the epilogue warps fold every
  // finished partial into the running sum M with round-to-nearest fp32 adds (tcgen05.ld / add / tcgen05.st) while
  // the MMA warp fills the other slot.  M never sees a tensor-core accumulation; the last partial of a tile is
  // added on the fly by the epilogue proper.
  // 3xTF32 keeps the split A operand in tensor memory as well: two slots (ping-pong between the splitter warps and the
  // MMA warp) of 64 columns — [A_hi: 32 K elements | A_lo: 32 K elements] for the 128 rows of the tile — after the
  // three accumulator regions.  The MMAs read A from there (tcgen05.mma with a TMEM A operand), so the shared-memory
  // pipe only carries the raw A tile once (TMA in, splitter out) and the B tiles: 112 KB per K iteration instead of
  // 192 KB — the kernel was bound by exactly that pipe (ncu: 59 % tensor-active, 1470 cycles per K iteration against
  // 768 of MMA work; 192 KB / 128 B per cycle = 1536).
  constexpr int kGroup = 8;
  constexpr int kTmemCols = X3 ? 512 : (2 * BN < 32 ? 32 : 2 * BN);
  constexpr int kASlotCols = 64;
  static_assert(!X3 || 3 * BN + 2 * kASlotCols <= 512, "TMEM budget of the 3xTF32 kernel");
  if (warp == 1) { if (CTA2) tmem_alloc2(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols); }
  tc_fence_before();
  if (CL) cluster_sync_all(); else __syncthreads();         // barriers of both CTAs initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // (warp-uniform for the compiler as well)

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = unit; t < total_tiles; t += n_units) {
        int n_tile, n0, oh0, ow0;
        tile_of(t, n_tile, n0, oh0, ow0);
        if (p.prefetch_side) {
          // pull the tile of the residual / addend / mask tensors towards L2 while the MMAs of this tile run, so
          // the epilogue's side reads are L2 hits
          const int cols_here = min(BN, p.Cout - n_tile * BN);
          for (int c = 0; c < cols_here; c += 32) {
            if (p.extra) tma_prefetch_l2_4d(&map_e, n_tile * BN + c, ow0, oh0, n0);
            if (p.mask) tma_prefetch_l2_4d(&map_m, n_tile * BN + c, ow0, oh0, n0);
          }
        }
        const int brow = n_tile * BN + (CTA2 ? (int)rank * kBRows : 0);     // first B row (output column) of this CTA
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % L::kStages;
          const uint32_t ph = (it / L::kStages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          const int tap = kit / p.cblocks, cb = kit - tap * p.cblocks;
          const int kh = tap / p.KW, kw = tap - kh * p.KW;
          uint8_t* sa = smem + s * L::kStageBytes;
          uint8_t* sb = sa + kBOff;
          const int kcol = tap * p.Cin + cb * BKE;
          if constexpr (MC) {
            // own A tile; our half of the B tile (rows [rank*BN/2, +BN/2)) goes to both CTAs, the other half comes
            // from the peer: every CTA's barrier still sees the bytes of one whole stage
            mbar_expect_tx(full_bar + s, L::kABytes + (X3 ? 2 : 1) * L::kBBytes);
            if (X3 && p.flat_w) tma_load_4d(&map_a, full_bar + s, sa, cb * BKE, ow0 + (kh - p.pad) * p.flat_w + kw - p.pad, 0, 0);
            else tma_load_4d(&map_a, full_bar + s, sa, cb * BKE, ow0 + kw - p.pad, oh0 + kh - p.pad, n0);
            const int half = (int)rank * (BN / 2);
            tma_load_2d_mc(&map_b, full_bar + s, sb + half * BKB, kcol, brow + half, (uint16_t)3);
            if (X3) tma_load_2d_mc(&map_b, full_bar + s, sb + L::kBBytes + half * BKB, kcol, p.b_lo_row + brow + half, (uint16_t)3);
          } else if constexpr (!CTA2) {
            mbar_expect_tx(full_bar + s, L::kABytes + (X3 ? 2 : 1) * L::kBBytes);
            if (p.stem) tma_load_5d(&map_a, full_bar + s, sa, 0, ow0, oh0 + (tap >> 1), tap & 1, n0);
            else if (X3 && p.flat_w) tma_load_4d(&map_a, full_bar + s, sa, cb * BKE, ow0 + (kh - p.pad) * p.flat_w + kw - p.pad, 0, 0);
            else tma_load_4d(&map_a, full_bar + s, sa, cb * BKE, ow0 + kw - p.pad, oh0 + kh - p.pad, n0);
            tma_load_2d(&map_b, full_bar + s, sb, kcol, brow);
            if (X3) tma_load_2d(&map_b, full_bar + s, sb + L::kBBytes, kcol, p.b_lo_row + brow);
          } else if constexpr (!X3) {
            // one barrier at the leader for the A tiles and B halves of both CTAs
            if (leader) mbar_expect_tx(full_bar + s, 2 * (L::kABytes + L::kBBytes));
            tma2_load_4d(&map_a, full_bar + s, sa, cb * BKE, ow0 + kw - p.pad, oh0 + kh - p.pad, n0);
            tma2_load_2d(&map_b, full_bar + s, sb, kcol, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (CTA pair: the leader only) ================================
    constexpr uint32_t idesc = make_idesc_tf32(CTA2 ? 2 * BM : BM, BN);
    // all 32 lanes run this loop converged; one elected lane issues each tcgen05 instruction (see umma_*_e)
    auto commit = [&](uint64_t* bar) { if constexpr (CTA2) umma_commit2_e(bar); else umma_commit_e(bar); };
    // release of a smem stage: MC tells both CTAs (each waits for two such arrivals before reloading the stage)
    auto commit_stage = [&](uint64_t* bar) {
      if constexpr (MC) umma_commit_mc_e(bar, (uint16_t)3); else commit(bar);
    };
    const uint32_t smem0 = smem_u32(smem);
    uint32_t it = 0;
    int local = 0;
    if (!leader) {
      // the peer's MMA warp only owns its TMEM allocation
    } else if constexpr (X3) {
      uint32_t grp = 0;                                      // partial-slot groups, counted across tiles
      const uint32_t a_tmem = tmem_base + (uint32_t)(3 * BN);
      for (int t = unit; t < total_tiles; t += n_units) {
        for (int g0 = 0; g0 < k_iters; g0 += kGroup, ++grp) {
          const int pp = grp & 1;
          mbar_wait(tmem_empty_bar + pp, ((grp >> 1) & 1) ^ 1);    // the folding warps have drained this slot
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)((1 + pp) * BN);
          const int g1 = min(g0 + kGroup, k_iters);
          for (int kit = g0; kit < g1; ++kit, ++it) {
            const int s = it % L::kStages;
            const int slot = it & 1;
            // the splitter warps (which waited for the stage's TMA bytes, B included) have written this A slot
            mbar_wait(split_bar + slot, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t b_hi = desc_lo_kmajor(smem0 + s * L::kStageBytes + kBOff);
            const uint32_t b_lo = desc_lo_kmajor(smem0 + s * L::kStageBytes + kBOff + L::kBBytes);
            const uint32_t a_hi = a_tmem + (uint32_t)(slot * kASlotCols);      // (A_lo sits 32 columns further)
            // 12 MMAs (4 K steps x {A_lo B_hi, A_hi B_lo, A_hi B_hi}), then the commits that free the smem stage
            // (B tiles, raw A tile) and the TMEM A slot
            umma_x3_kiter<MC>(d, a_hi, b_hi, b_lo, idesc, kit > g0 ? 1u : 0u, empty_bar + s, aslot_bar + slot);
            if (kit == g1 - 1) commit(tmem_full_bar + pp);                 // partial complete
          }
        }
      }
    } else {
    for (int t = unit; t < total_tiles; t += n_units, ++local) {
      const int acc = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      // the epilogue (of both CTAs) has drained this accumulator
      if (CTA2) mbar_wait_cluster(tmem_empty_bar + acc, aph ^ 1); else mbar_wait(tmem_empty_bar + acc, aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kit = 0; kit < k_iters; ++kit, ++it) {
        const int s = it % L::kStages;
        const uint32_t ph = (it / L::kStages) & 1;
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        const uint32_t a_d = desc_lo_kmajor(smem0 + s * L::kStageBytes);
        const uint32_t b_d = desc_lo_kmajor(smem0 + s * L::kStageBytes + kBOff);
        // 4 MMAs (K steps of 8) and the commit that frees the smem stage when they retire
        umma_tf32_kiter<PAIR>(tmem_d, a_d, b_d, idesc, kit ? 1u : 0u, empty_bar + s);
        if (kit == k_iters - 1) commit(tmem_full_bar + acc);  // accumulator complete
      }
    }
    }
  } else if (X3 && warp >= 6) {
    // ================================ operand splitter (warps 6..9, 3xTF32 mode) ================================
    // Each thread owns one row of the A tile (= one TMEM lane): it reads the row's 32 fp32 activations from the landed
    // (128B-swizzled) tile, splits every element into its TF32-exact high part (low 13 mantissa bits cleared) and the
    // TF32-exact part of the remainder, and stores both into the TMEM A slot [hi: 32 columns | lo: 32 columns].  Both
    // parts are exactly representable in TF32, so the result does not depend on how the tensor core converts fp32
    // bit patterns.  Nothing is written back to shared memory.
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    const uint32_t a_tmem = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(3 * BN);
    // Two instances of the loop (the chain TMA -> splitter -> MMA is latency-critical: the tap test of the dense-axis
    // 3x3 mode in the common instance cost the RPN conv 2-8 %): FLAT carries the per-row tap mask, the other nothing.
    auto split_loop = [&](auto flat_tag) {
      constexpr bool FLAT = decltype(flat_tag)::value;
      uint32_t it = 0;
      for (int t = unit; t < total_tiles; t += n_units) {
        // dense-axis 3x3 mode: bit `tap` set = that tap of this row's pixel lies inside the pixel's own map
        uint32_t tapmask = 0xFFFFFFFFu;
        if constexpr (FLAT) {
          int n_tile, n0, oh0, ow0;
          tile_of(t, n_tile, n0, oh0, ow0);
          const int rem = (ow0 + row) % (p.flat_h * p.flat_w);
          const int oh = rem / p.flat_w, ow = rem - oh * p.flat_w;
          tapmask = 0u;
          for (int tap = 0; tap < p.taps; ++tap) {
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            if ((unsigned)(oh + kh - p.pad) < (unsigned)p.flat_h && (unsigned)(ow + kw - p.pad) < (unsigned)p.flat_w)
              tapmask |= 1u << tap;
          }
        }
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % L::kStages;
          const uint32_t ph = (it / L::kStages) & 1;
          const int slot = it & 1;
          mbar_wait(full_bar + s, ph);                                  // the stage's bytes have landed
          mbar_wait(aslot_bar + slot, ((it >> 1) & 1) ^ 1);             // the MMAs of two K-iterations ago have read the slot
          tc_fence_after();
          const uint32_t a_row = smem_u32(smem + s * L::kStageBytes) + row_off;
          uint32_t hi[32], lo[32];
          bool zero_row = false;
          if constexpr (FLAT) zero_row = !((tapmask >> (kit / p.cblocks)) & 1u);
          if (zero_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { hi[j] = 0u; lo[j] = 0u; }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {                  // quarter-warp phases hit 8 distinct 16-byte bank groups
              float v[4];
              lds128(a_row + ((j ^ sw) << 4), v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t h = __float_as_uint(v[e]) & 0xFFFFE000u;
                hi[4 * j + e] = h;
                lo[4 * j + e] = __float_as_uint(v[e] - __uint_as_float(h)) & 0xFFFFE000u;
              }
            }
          }
          tmem_st32(a_tmem + (uint32_t)(slot * kASlotCols), hi);
          tmem_st32(a_tmem + (uint32_t)(slot * kASlotCols + 32), lo);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(split_bar + slot);
        }
      }
    };
    if (p.flat_w) split_loop(std::true_type{}); else split_loop(std::false_type{});
  } else {
    // ================================ epilogue (warps 2..5) ================================
    // Each warp owns TMEM lanes / tile rows [32q, 32q+32) end to end (its own staging slices, its own TMA
    // stores), so the chunk loop needs only __syncwarp.  EPI selects the compiled side-input handling.
    constexpr bool kExtra = (EPI & EPI_EXTRA) != 0, kMask = (EPI & EPI_MASK) != 0, kScalar = (EPI & EPI_SCALAR) != 0;
    constexpr bool kAhead = !kScalar && (kExtra != kMask);          // exactly one side input: prefetch it a chunk ahead
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // tile row owned in the TMEM -> smem pass
    const int pc = lane & 7;                             // re-mapped pass: 16-byte column group of the chunk,
    const int pr = lane >> 3;                            //   rows pr + 4 i of the warp's 32
    const int wl = row % p.tw, hl = (row / p.tw) % p.th, nl = row / (p.tw * p.th);
    const int r0 = quarter * 32;                         // the warp's rows as a sub-box of the tile
    const int sub_w = r0 % p.tw, sub_h = (r0 / p.tw) % p.th, sub_n = r0 / (p.tw * p.th);
    const uint32_t my_stage = smem_u32(staging) + quarter * 4096;      // + (chunk & 1) * kStagingBytes
    const uint32_t my_rows = smem_u32(row_pix) + quarter * 128;
    const uint32_t st_off = lane * 128;                  // TMEM -> smem pass: own row, chunk j at (j ^ (lane & 7)) * 16
    const uint32_t st_sw = lane & 7;
    const float lo = p.relu ? 0.f : -INFINITY;
    // hand a TMEM accumulator / partial slot back to the (leader's) MMA warp
    auto release_tmem = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2 && !leader) mbar_arrive_remote(bar, 0); else mbar_arrive(bar);
      }
    };
    uint32_t chunk_ctr = 0;
    uint32_t x3_grp = 0;                                   // 3xTF32: partial-slot groups, counted across tiles
    int local = 0;
    for (int t = unit; t < total_tiles; t += n_units, ++local) {
      const int acc = X3 ? 0 : (local & 1);
      const uint32_t aph = (local >> 1) & 1;
      int n_tile, n0, oh0, ow0;
      tile_of(t, n_tile, n0, oh0, ow0);
      int pix[8];
      if (kExtra || kMask || kScalar) {                  // side reads / scalar stores need the rows' pixel index
        __syncwarp();                                    // lanes are done reading the previous tile's row table
        const int n = n0 + nl, oh = oh0 + hl, ow = ow0 + wl;
        const bool ok = n < p.N && oh < p.OH && ow < p.OW;
        sts32(my_rows + lane * 4, ok ? (n * p.out_H + oh * p.os) * p.out_W + ow * p.os : -1);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) pix[i] = lds32(my_rows + (pr + 4 * i) * 4);
      }
      const int cols_here = min(BN, p.Cout - n_tile * BN);
      const int n_chunks = (cols_here + 31) >> 5;
      const int colbase = n_tile * BN + pc * 4;
      float sc[4], bi[4];                                // per-channel scale / bias, fetched one chunk ahead
      load_scale_bias(p, colbase, sc, bi);
      const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);

      // side reads (residual / addend and ReLU mask) of chunk `ch`, all in flight at once; issued ONE CHUNK AHEAD
      // of their use when a single side input is present (two register sets), else right before use
      auto issue_side = [&](int ch, float4 (&ex)[8], float4 (&mk)[8]) {
        const int col = colbase + ch * 32;
        const bool col_ok = col < p.Cout;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = pix[i] >= 0 && col_ok;
          const size_t off = ok ? (size_t)pix[i] * p.ldc + col : 0;
          if (kExtra) ex[i] = ok ? dd::ldg4(p.extra + off) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (kMask) mk[i] = ok ? dd::ldg4(p.mask + off) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
      };

      auto process = [&](const uint32_t (&r)[32], int ch, float4 (&ex)[8], float4 (&mk)[8]) {
        const uint32_t stg = my_stage + (chunk_ctr & 1) * L::kStagingBytes;
        if (!kScalar) {
          if (lane == 0) tma_store_wait_read<1>();       // the store that last read this slice has drained
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(stg + st_off + ((j ^ st_sw) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                 __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        // ---- re-mapped pass: lane = (column group pc, rows pr + 4 i)
        const int col = colbase + ch * 32;
        const bool col_ok = col < p.Cout;
        if (!kScalar && !kAhead && (kExtra || kMask)) issue_side(ch, ex, mk);
        float scn[4], bin[4];
        if (ch + 1 < n_chunks) load_scale_bias(p, col + 32, scn, bin);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = pr + 4 * i;
          const uint32_t sp = stg + rl * 128 + ((pc ^ (rl & 7)) << 4);
          float v[4];
          lds128(sp, v);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = fmaf(v[e], sc[e], bi[e]);
          if (!kScalar) {
            if (kExtra) { v[0] += ex[i].x; v[1] += ex[i].y; v[2] += ex[i].z; v[3] += ex[i].w; }
            if (kMask) {
              v[0] = mk[i].x > 0.f ? v[0] : 0.f; v[1] = mk[i].y > 0.f ? v[1] : 0.f;
              v[2] = mk[i].z > 0.f ? v[2] : 0.f; v[3] = mk[i].w > 0.f ? v[3] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], lo);
            sts128(sp, v[0], v[1], v[2], v[3]);
          } else if (pix[i] >= 0 && col_ok) {            // narrow / unaligned outputs: scalar guarded path
            const size_t off = (size_t)pix[i] * p.ldc + col;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (col + e < p.Cout) {
                float x = v[e];
                if (p.extra) x += __ldg(p.extra + off + e);
                if (p.mask) x = __ldg(p.mask + off + e) > 0.f ? x : 0.f;
                p.out[off + e] = fmaxf(x, lo);
              }
            }
          }
        }
        if (ch + 1 < n_chunks) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { sc[e] = scn[e]; bi[e] = bin[e]; }
        }
        if (!kScalar) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_4d(&map_c, stg, n_tile * BN + ch * 32, ow0 + sub_w, oh0 + sub_h, n0 + sub_n);
        } else {
          __syncwarp();
        }
        ++chunk_ctr;
      };

      uint32_t ra[32], rb[32];
      if constexpr (kOnePass) {
        // ---- one-pass epilogue (3xTF32 + residual, short K: the memory-bound 1x1 layers of the trunk).  The two-pass
        // loop below is issue-bound on its four warps (~350 instructions per chunk and lane: transpose through the
        // staging buffer, address arithmetic of the coalesced side reads; ncu: 6.6 us per tile on res2.conv3 for 3.3
        // of data movement).  Here the warp's [32 rows x 32 ch] residual sub-box is landed by TMA in the SAME swizzled
        // row layout as the staging buffer (two slices, loads issued two chunks ahead), the per-channel scale / bias
        // of the tile sit in a small table read by broadcast, and every lane finishes its own row straight out of the
        // accumulator registers: tcgen05.ld -> fma, + residual, ReLU -> staging -> TMA store.  Same arithmetic order.
        const uint32_t my_res = smem_u32(smem + L::kResOff) + quarter * 8192;
        const uint32_t my_tab = smem_u32(smem + L::kTabOff) + quarter * 1024;         // [BN scale | BN bias]
        uint64_t* my_bar = res_bar + quarter * 2;
        auto issue_res = [&](uint32_t c, int ch) {       // chunk counter c (buffer c & 1), chunk ch of this tile
          if (lane == 0) {
            mbar_expect_tx(my_bar + (c & 1), 4096);
            tma_load_4d(&map_e, my_bar + (c & 1), smem + L::kResOff + quarter * 8192 + (c & 1) * 4096,
                        n_tile * BN + ch * 32, ow0 + sub_w, oh0 + sub_h, n0 + sub_n);
          }
        };
        if (kExtra) {
          issue_res(chunk_ctr, 0);
          if (n_chunks > 1) issue_res(chunk_ctr + 1, 1);
        }
        {
          float s4[4], b4[4];
          if (lane * 4 < BN) {
            load_scale_bias(p, n_tile * BN + lane * 4, s4, b4);
            sts128(my_tab + lane * 16, s4[0], s4[1], s4[2], s4[3]);
            sts128(my_tab + BN * 4 + lane * 16, b4[0], b4[1], b4[2], b4[3]);
          }
          __syncwarp();
        }
        const int n_groups = (k_iters + kGroup - 1) / kGroup;
        for (int g = 0; g < n_groups; ++g, ++x3_grp) {
          const int pp = x3_grp & 1;
          mbar_wait(tmem_full_bar + pp, (x3_grp >> 1) & 1);
          tc_fence_after();
          const uint32_t tp = tm + (uint32_t)((1 + pp) * BN);
#pragma unroll 1
          for (int ch = 0; ch < n_chunks; ++ch) {
            tmem_ld32(tp + (uint32_t)(ch * 32), ra);
            if (g > 0) {
              tmem_ld32(tm + (uint32_t)(ch * 32), rb);
#pragma unroll
              for (int j = 0; j < 32; ++j) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(rb[j]));
            }
            if (g < n_groups - 1) { tmem_st32(tm + (uint32_t)(ch * 32), ra); continue; }
            if (ch == n_chunks - 1) release_tmem(tmem_empty_bar + pp);     // slot fully read: hand it back
            const uint32_t stg = my_stage + (chunk_ctr & 1) * L::kStagingBytes;
            if (lane == 0) tma_store_wait_read<1>();     // the store that last read this staging slice has drained
            if (kExtra) mbar_wait(my_bar + (chunk_ctr & 1), (chunk_ctr >> 1) & 1);       // the residual slice has landed
            __syncwarp();
            const uint32_t res = my_res + (chunk_ctr & 1) * 4096 + st_off;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float s4[4], b4[4], e4[4] = {0.f, 0.f, 0.f, 0.f};
              lds128(my_tab + (ch * 32 + j * 4) * 4, s4);
              lds128(my_tab + BN * 4 + (ch * 32 + j * 4) * 4, b4);
              if (kExtra) lds128(res + ((j ^ st_sw) << 4), e4);
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[e] = fmaf(__uint_as_float(ra[4 * j + e]), s4[e], b4[e]);
                if (kExtra) v[e] += e4[e];
                v[e] = fmaxf(v[e], lo);
              }
              sts128(stg + st_off + ((j ^ st_sw) << 4), v[0], v[1], v[2], v[3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_4d(&map_c, stg, n_tile * BN + ch * 32, ow0 + sub_w, oh0 + sub_h, n0 + sub_n);
            if (kExtra && ch + 2 < n_chunks) issue_res(chunk_ctr + 2, ch + 2);       // (every lane has read the slice)
            ++chunk_ctr;
          }
          if (g < n_groups - 1) release_tmem(tmem_empty_bar + pp);
        }
        continue;
      }
      float4 exa[8], mka[8], exb[kAhead ? 8 : 1], mkb[kAhead ? 8 : 1];
      if (kAhead) issue_side(0, exa, mka);             // does not depend on the accumulator: before the wait
      if constexpr (X3) {
        // fold the finished partial slots into M (round-to-nearest adds); the last partial of the tile is added
        // on the fly in the chunk loop below.  tm = TMEM address of M for this warp's lanes.
        const int n_groups = (k_iters + kGroup - 1) / kGroup;
        for (int g = 0; g < n_groups; ++g, ++x3_grp) {
          const int pp = x3_grp & 1;
          mbar_wait(tmem_full_bar + pp, (x3_grp >> 1) & 1);
          tc_fence_after();
          const uint32_t tp = tm + (uint32_t)((1 + pp) * BN);
          if (g < n_groups - 1) {
#pragma unroll 1
            for (int ch = 0; ch < n_chunks; ++ch) {
              tmem_ld32(tp + (uint32_t)(ch * 32), ra);
              if (g > 0) {
                tmem_ld32(tm + (uint32_t)(ch * 32), rb);
#pragma unroll
                for (int j = 0; j < 32; ++j) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(rb[j]));
              }
              tmem_st32(tm + (uint32_t)(ch * 32), ra);
            }
          } else {
            // the last partial: added on the fly, then the epilogue proper.  The side input (residual) of chunk ch + 1
            // is requested before chunk ch is processed (two register sets), like in the TF32 loop below: requested
            // right before its use, every chunk paid the full DRAM latency (res2.conv3: 6.6 us per tile)
            auto fetch = [&](int ch) {
              tmem_ld32(tp + (uint32_t)(ch * 32), ra);
              if (g > 0) {
                tmem_ld32(tm + (uint32_t)(ch * 32), rb);
#pragma unroll
                for (int j = 0; j < 32; ++j) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(rb[j]));
              }
              if (ch == n_chunks - 1) release_tmem(tmem_empty_bar + pp);   // slot fully read: hand it back before the slow part
            };
#pragma unroll 1
            for (int ch = 0; ch < n_chunks; ch += 2) {
              fetch(ch);
              if constexpr (kAhead) { if (ch + 1 < n_chunks) issue_side(ch + 1, exb, mkb); }
              process(ra, ch, exa, mka);
              if (ch + 1 < n_chunks) {
                fetch(ch + 1);
                if constexpr (kAhead) {
                  if (ch + 2 < n_chunks) issue_side(ch + 2, exa, mka);
                  process(ra, ch + 1, exb, mkb);
                } else {
                  process(ra, ch + 1, exa, mka);
                }
              }
            }
            continue;
          }
          release_tmem(tmem_empty_bar + pp);
        }
        continue;
      }
      mbar_wait(tmem_full_bar + acc, aph);
      tc_fence_after();
      tmem_ld32_nowait(tm, ra);
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ch += 2) {
        tmem_ld_wait();
        if (ch + 1 < n_chunks) {
          tmem_ld32_nowait(tm + (uint32_t)((ch + 1) * 32), rb);   // moves while ra is processed
          if constexpr (kAhead) issue_side(ch + 1, exb, mkb);
        }
        process(ra, ch, exa, mka);
        if (ch + 1 < n_chunks) {
          tmem_ld_wait();
          if (ch + 2 < n_chunks) {
            tmem_ld32_nowait(tm + (uint32_t)((ch + 2) * 32), ra);
            if (kAhead) issue_side(ch + 2, exa, mka);
          }
          if constexpr (kAhead) process(rb, ch + 1, exb, mkb);
          else process(rb, ch + 1, exa, mka);
        }
      }
      // every tcgen05.ld of this accumulator has completed: hand it back to the MMA warp
      release_tmem(tmem_empty_bar + acc);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  if (CL) cluster_sync_all(); else __syncthreads();     // (pair: the peer may still be arriving on our barriers)
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------ wgrad kernel
// D[co, ci] (one filter tap per CTA column) = sum over a range of pixel blocks of GY[pix, co] * X[pix + tap, ci].
// Both operands are MN-major: a [32 pixels x 32 channels] block with the 32 channels (128 B) contiguous is one
// UMMA "MN-major, 128B-swizzle" column block; a stage holds 4 (co) + BN/32 (ci) of them, each operand landed by
// ONE 5-D TMA box whose outermost dimension walks the 32-channel blocks.  1x1 convs see the pixels as one
// dense axis (no padding of 7x7 ROI maps to 8x8).
struct TcWgradParams {
  float* gw;               // [Cout][taps*Cin], accumulated into
  const float* scale;      // per-Cout BN scale (may be null)
  int Cout, Cin, taps, KW, pad;
  int tn, th, tw;          // pixel block = tn*th*tw == 32
  int tiles_h, tiles_w, pix_blocks, blocks_per_split;
  int ci_tiles, co_tiles, splits;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int WG_KPIX = 32;                       // pixels per stage (4 MMA K-steps of 8)
constexpr int WG_BOX_BYTES = WG_KPIX * 128;       // one [32 pix x 32 ch] box = 4 KB

// MN-major TF32 operands exist only in the "128B swizzle with 32B atom" layout (cute::UMMA::LayoutType::
// SWIZZLE_128B_BASE32B = 1; CUTLASS sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout"), produced by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: rows of 128 B
// (32 channels), 32-byte chunks XOR-swizzled with (row & 3), pattern period 4 rows = 512 B.
// LBO = byte distance between 32-channel column blocks, SBO = between 4-pixel row groups.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(WG_BOX_BYTES >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int m, int n) {
  return make_idesc_tf32(m, n) | (1u << 15) | (1u << 16);
}

// X3 (3xTF32): a stage holds [A_hi | A_lo | B_hi | B_lo]; both operands are activations and are split in the kernel.
template <int BN, bool X3>
struct WgradSmem {
  static constexpr int kStages = X3 ? (BN >= 128 ? 3 : 4) : (BN == 256 ? 4 : 6);
  static constexpr int kABytes = 4 * WG_BOX_BYTES;            // 128 co
  static constexpr int kBBytes = (BN / 32) * WG_BOX_BYTES;
  static constexpr int kStageBytes = (X3 ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kTotal = kStages * kStageBytes + 1024 + 256;
};

// Persistent: grid = min(#work items, #SMs); a work item = (filter tap, ci tile, co tile, pixel split); the items
// are walked in a static round-robin by every role.  Two TMEM accumulators (ping-pong) let the L2 reductions of
// item i overlap the main loop of item i+1; the TMA producer runs ahead across items through the smem ring.
template <int BN, bool X3>
__global__ void __launch_bounds__(X3 ? NUM_THREADS_X3 : NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
                const TcWgradParams p) {
  using L = WgradSmem<BN, X3>;
  constexpr int kALo = X3 ? L::kABytes : 0;                         // A_lo right after A_hi
  constexpr int kBOff = (X3 ? 2 : 1) * L::kABytes;                  // B_hi
  constexpr int kBLo = X3 ? L::kBBytes : 0;                         // B_lo right after B_hi
  constexpr int kAccCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kStages * L::kStageBytes);
  uint64_t* empty_bar = full_bar + L::kStages;
  uint64_t* tmem_full_bar = empty_bar + L::kStages;                 // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                     // [2]
  uint64_t* split_bar = tmem_empty_bar + 2;                         // [kStages], X3 only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(split_bar + L::kStages);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (warp-uniform: see conv_tc_kernel)
  const int total_items = p.taps * p.ci_tiles * p.co_tiles * p.splits;

  // item -> (tap, ci tile, co tile, split): consecutive items share the pixel range (L2 reuse of gy / x tiles)
  auto decode = [&](int w, int& tap, int& ci0, int& co0, int& pb_begin, int& k_iters) {
    const int xi = w % (p.taps * p.ci_tiles); w /= (p.taps * p.ci_tiles);
    const int co_tile = w % p.co_tiles;
    const int split = w / p.co_tiles;
    tap = xi / p.ci_tiles;
    ci0 = (xi % p.ci_tiles) * BN;
    co0 = co_tile * BM;
    pb_begin = split * p.blocks_per_split;
    k_iters = max(min(p.pix_blocks, pb_begin + p.blocks_per_split) - pb_begin, 0);
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_gy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < L::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    if (X3) for (int s = 0; s < L::kStages; ++s) mbar_init(split_bar + s, 128);
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar + s, 1); mbar_init(tmem_empty_bar + s, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        int tap, ci0, co0, pb_begin, k_iters;
        decode(w, tap, ci0, co0, pb_begin, k_iters);
        const int kh = tap / p.KW, kw = tap % p.KW;
        for (int kit = 0; kit < k_iters; ++kit, ++it) {
          const int s = it % L::kStages;
          const uint32_t ph = (it / L::kStages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          int pb = pb_begin + kit;
          const int bw = pb % p.tiles_w; pb /= p.tiles_w;
          const int bh = pb % p.tiles_h; pb /= p.tiles_h;
          const int n0 = pb * p.tn, oh0 = bh * p.th, ow0 = bw * p.tw;
          uint8_t* sa = smem + s * L::kStageBytes;
          uint8_t* sb = sa + kBOff;
          mbar_expect_tx(full_bar + s, L::kABytes + L::kBBytes);
          // 5-D views {32 ch, W, H, N, channel block}: one instruction lands all [32 pix x 32 ch] column blocks
          tma_load_5d(&map_gy, full_bar + s, sa, 0, ow0, oh0, n0, co0 / 32);
          tma_load_5d(&map_x, full_bar + s, sb, 0, ow0 + kw - p.pad, oh0 + kh - p.pad, n0, ci0 / 32);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32_mn(BM, BN);
    uint32_t it = 0;
    int local = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      int tap, ci0, co0, pb_begin, k_iters;
      decode(w, tap, ci0, co0, pb_begin, k_iters);
      if (k_iters == 0) continue;                         // (empty split: nothing to accumulate or reduce)
      const int acc = local & 1;
      mbar_wait(tmem_empty_bar + acc, ((local >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * kAccCols);
      for (int kit = 0; kit < k_iters; ++kit, ++it) {
        const int s = it % L::kStages;
        const uint32_t ph = (it / L::kStages) & 1;
        mbar_wait((X3 ? split_bar : full_bar) + s, ph);
        tc_fence_after();
        {     // all lanes converged; one elected lane issues (see the note at umma_tf32_ss_e)
          const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
          const uint32_t b_addr = a_addr + kBOff;
#pragma unroll
          for (int k = 0; k < WG_KPIX / UMMA_K; ++k) {
            const uint64_t ad = make_mnmajor_sw128_desc(a_addr + k * 1024);
            const uint64_t bd = make_mnmajor_sw128_desc(b_addr + k * 1024);
            if (X3) {                                         // small terms first
              const uint64_t adl = make_mnmajor_sw128_desc(a_addr + kALo + k * 1024);
              const uint64_t bdl = make_mnmajor_sw128_desc(b_addr + kBLo + k * 1024);
              umma_tf32_e(d, adl, bd, idesc, (kit | k) ? 1u : 0u);
              umma_tf32_e(d, ad, bdl, idesc, 1u);
              umma_tf32_e(d, ad, bd, idesc, 1u);
            } else {
              umma_tf32_e(d, ad, bd, idesc, (kit | k) ? 1u : 0u);
            }
          }
          umma_commit_e(empty_bar + s);
          if (kit == k_iters - 1) umma_commit_e(tmem_full_bar + acc);
        }
      }
      ++local;
    }
  } else if (X3 && warp >= 6) {
    // operand splitter (3xTF32): both landed tiles -> TF32-exact high parts in place, low parts next to them
    const int rs = threadIdx.x - 192;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      int tap, ci0, co0, pb_begin, k_iters;
      decode(w, tap, ci0, co0, pb_begin, k_iters);
      for (int kit = 0; kit < k_iters; ++kit, ++it) {
        const int s = it % L::kStages;
        const uint32_t ph = (it / L::kStages) & 1;
        mbar_wait(full_bar + s, ph);
        const uint32_t base = smem_u32(smem + s * L::kStageBytes) + rs * 16;
#pragma unroll 4
        for (int j = 0; j < (L::kABytes + L::kBBytes) / 2048; ++j) {
          const bool is_b = j * 2048 >= L::kABytes;
          const uint32_t hi_addr = base + (is_b ? kBOff + (j * 2048 - L::kABytes) : j * 2048);
          const uint32_t lo_addr = hi_addr + (is_b ? kBLo : kALo);
          float v[4];
          lds128(hi_addr, v);
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xFFFFE000u);
            lo[e] = __uint_as_float(__float_as_uint(v[e] - hi[e]) & 0xFFFFE000u);
          }
          sts128(hi_addr, hi[0], hi[1], hi[2], hi[3]);
          sts128(lo_addr, lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        mbar_arrive(split_bar + s);
      }
    }
  } else {
    // epilogue: BN scale, then fp32 vector reductions straight into gw (split-K partial sums meet in L2; no
    // partial buffers, no reduce pass)
    const int quarter = warp & 3;
    const size_t ldp = (size_t)p.taps * p.Cin;
    int local = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      int tap, ci0, co0, pb_begin, k_iters;
      decode(w, tap, ci0, co0, pb_begin, k_iters);
      if (k_iters == 0) continue;
      const int acc = local & 1;
      const int co = co0 + quarter * 32 + lane;
      float* dst_row = p.gw + (size_t)co * ldp + (size_t)tap * p.Cin + ci0;
      const float sc = (p.scale && co < p.Cout) ? __ldg(p.scale + co) : 1.f;
      mbar_wait(tmem_full_bar + acc, (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tm + (uint32_t)c0, r);
        if (c0 + 32 >= BN) {                              // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(tmem_empty_bar + acc);
        }
        if (co < p.Cout && ci0 + c0 < p.Cin) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst_row + c0 + j, __uint_as_float(r[j]) * sc, __uint_as_float(r[j + 1]) * sc,
                       __uint_as_float(r[j + 2]) * sc, __uint_as_float(r[j + 3]) * sc);
        }
      }
      ++local;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kAccCols);
  }
}

// W'[ci][KH-1-kh][KW-1-kw][co] = scale[co] * W[co][kh][kw][ci]  (dgrad as a forward conv over GY).
// One 32x32 smem-tiled transpose per (tap, co block, ci block): coalesced on both sides.
// lo_off > 0 (3xTF32 mode): wt receives the TF32-exact high parts and wt + lo_off the low parts.
__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                             float* __restrict__ wt, int Cout, int Cin, int taps, long long lo_off) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z, co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int co = co0 + ty + j, ci = ci0 + tx;
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      v = w[((size_t)co * taps + tap) * Cin + ci];
      if (scale) v *= __ldg(scale + co);
    }
    tile[ty + j][tx] = v;
  }
  __syncthreads();
  const int tflip = taps - 1 - tap;                  // (KH-1-kh)*KW + (KW-1-kw)
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int ci = ci0 + ty + j, co = co0 + tx;
    if (co < Cout && ci < Cin) {
      const float v = tile[tx][ty + j];
      const size_t o = ((size_t)ci * taps + tflip) * Cout + co;
      if (lo_off > 0) {
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        wt[o] = hi;
        wt[o + lo_off] = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
      } else {
        wt[o] = v;
      }
    }
  }
}

// The same for up to 16 layers in ONE launch (a ResNet stage prepares all its dgrad weights at once): blockIdx.x
// walks the concatenated tile lists of the entries.
struct PrepEntry {
  const float* w; const float* scale; float* wt;
  int Cout, Cin, taps, tile_begin;
  long long lo_off;
};
struct PrepBatch { PrepEntry e[16]; int n; };

__global__ void weight_flip_transpose_batch_kernel(const PrepBatch b) {
  __shared__ float tile[32][33];
  int k = 0;
  while (k + 1 < b.n && (int)blockIdx.x >= b.e[k + 1].tile_begin) ++k;
  const PrepEntry& E = b.e[k];
  int t = blockIdx.x - E.tile_begin;
  const int ci_blocks = (E.Cin + 31) / 32, co_blocks = (E.Cout + 31) / 32;
  const int ci0 = (t % ci_blocks) * 32; t /= ci_blocks;
  const int co0 = (t % co_blocks) * 32;
  const int tap = t / co_blocks;
  const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int co = co0 + ty + j, ci = ci0 + tx;
    float v = 0.f;
    if (co < E.Cout && ci < E.Cin) {
      v = E.w[((size_t)co * E.taps + tap) * E.Cin + ci];
      if (E.scale) v *= __ldg(E.scale + co);
    }
    tile[ty + j][tx] = v;
  }
  __syncthreads();
  const int tflip = E.taps - 1 - tap;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int ci = ci0 + ty + j, co = co0 + tx;
    if (co < E.Cout && ci < E.Cin) {
      const float v = tile[tx][ty + j];
      const size_t o = ((size_t)ci * E.taps + tflip) * E.Cout + co;
      if (E.lo_off > 0) {
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        E.wt[o] = hi;
        E.wt[o + E.lo_off] = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
      } else {
        E.wt[o] = v;
      }
    }
  }
}

__global__ void fill_kernel(const float* __restrict__ addend, const float* __restrict__ mask, float* __restrict__ gx,
                            long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    float v = addend ? addend[t] : 0.f;
    if (mask) v = mask[t] > 0.f ? v : 0.f;
    gx[t] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host side
int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// cuTensorMapEncodeTiled is a driver-API entry point; it is resolved through the runtime
// (cudaGetDriverEntryPoint) so that the library has no link-time dependency on libcuda.so and still loads on
// a machine without a driver (the CPU build/ABI check).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return dd::fail(-1, "cuTensorMapEncodeTiled is unavailable (no CUDA driver)", __FILE__, __LINE__);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(dd::g_err, sizeof(dd::g_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return -1;
  }
  return 0;
}

// CTA-pair mode switch (default on).  DD_TC_CTA2=0 in the environment (read once) keeps every launch on single CTAs:
// used to A/B the two modes on the GPU.
bool tc_cta2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DD_TC_CTA2");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

int tc_cta2_forced() {         // 0 = heuristics, 1 = cta_group::2 pairs everywhere, 2 = multicast pairs everywhere
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DD_TC_CTA2");
    mode = (e != nullptr && e[0] == '2') ? 1 : ((e != nullptr && e[0] == '3') ? 2 : 0);
  }
  return mode;
}

template <int BN, int EPI, bool X3, int PAIR>
int launch_tc4(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& me,
               const CUtensorMap& mm, const TcParams& p, cudaStream_t s) {
  using L = SmemLayout<BN, X3, PAIR, (EPI & EPI_ONEPASS) != 0>;
  constexpr bool CL = PAIR != 0;
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, EPI, X3, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int m_units = CL ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const long long total = (long long)m_units * p.n_tiles;
  const int max_units = CL ? dd::sm_budget() / 2 : dd::sm_budget();
  const int units = total < max_units ? (int)total : max_units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CL ? 2 * units : units));
  cfg.blockDim = dim3(X3 ? NUM_THREADS_X3 : NUM_THREADS);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, EPI, X3, PAIR>, ma, mb, mc, me, mm, p));
  DD_LAUNCHED();
  return 0;
}

template <int BN, int EPI, bool X3>
int launch_tc3(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& me,
               const CUtensorMap& mm, const TcParams& p, int pair, cudaStream_t s) {
  if constexpr (!X3) {
    if (pair == 1) return launch_tc4<BN, EPI, X3, 1>(ma, mb, mc, me, mm, p, s);
  }
  if (pair == 2) return launch_tc4<BN, EPI, X3, 2>(ma, mb, mc, me, mm, p, s);
  return launch_tc4<BN, EPI, X3, 0>(ma, mb, mc, me, mm, p, s);
}

template <int BN, int EPI>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& me,
               const CUtensorMap& mm, const TcParams& p, bool x3, int pair, cudaStream_t s) {
  if constexpr (BN <= 128) {
    if (x3) return launch_tc3<BN, EPI, true>(ma, mb, mc, me, mm, p, pair, s);
  }
  if constexpr ((EPI & EPI_ONEPASS) != 0) return dd::fail(-1, "one-pass epilogue needs the 3xTF32 kernel", __FILE__, __LINE__);
  else return launch_tc3<BN, EPI, false>(ma, mb, mc, me, mm, p, pair, s);
}

template <int BN>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& me,
              const CUtensorMap& mm, const TcParams& p, bool x3, int pair, cudaStream_t s) {
  if (!p.tma_store) return launch_tc2<BN, EPI_SCALAR>(ma, mb, mc, me, mm, p, x3, pair, s);
  if constexpr (BN <= 128) {
    if (p.onepass && p.extra) return launch_tc2<BN, EPI_EXTRA | EPI_ONEPASS>(ma, mb, mc, me, mm, p, x3, pair, s);
    if (p.onepass) return launch_tc2<BN, EPI_ONEPASS>(ma, mb, mc, me, mm, p, x3, pair, s);
  }
  const int epi = (p.extra ? EPI_EXTRA : 0) | (p.mask ? EPI_MASK : 0);
  switch (epi) {
    case 0: return launch_tc2<BN, 0>(ma, mb, mc, me, mm, p, x3, pair, s);
    case EPI_EXTRA: return launch_tc2<BN, EPI_EXTRA>(ma, mb, mc, me, mm, p, x3, pair, s);
    case EPI_MASK: return launch_tc2<BN, EPI_MASK>(ma, mb, mc, me, mm, p, x3, pair, s);
    default: return launch_tc2<BN, EPI_EXTRA | EPI_MASK>(ma, mb, mc, me, mm, p, x3, pair, s);
  }
}

bool tc_onepass_enabled() {         // DD_TC_ONEPASS=0 keeps the two-pass epilogue everywhere (A/B runs)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DD_TC_ONEPASS");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// 3xTF32 runs tiles of at most 128 columns: TMEM holds the running sum and two partial slots of a tile
int tc_bn_for(int ncols, bool x3) {
  if (x3) return ncols > 64 ? 128 : 64;
  return (ncols % 256 == 0) ? 256 : (ncols > 64 ? 128 : 64);
}
// rows of one plane of a prepared (hi / lo split) B matrix: the column count rounded up to whole 3xTF32 tiles
int tc_rows_pad(int ncols) { return ncols > 64 ? (ncols + 127) / 128 * 128 : 64; }

// out[r][k] = hi(w[r][k]), out[rows_pad + r][k] = lo(w[r][k]); rows in [rows, rows_pad) are zero in both planes
__global__ void split_hi_lo_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, int rows_pad,
                                   long long K) {
  const long long total = (long long)rows_pad * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / K;
    const float v = r < rows ? w[t] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    out[t] = hi;
    out[total + t] = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
  }
}

// The same for up to 32 weight matrices in one launch (the forward weights of a whole model, prepared once per step).
struct SplitBatch {
  struct Entry { const float* w; float* out; int rows, rows_pad; long long K; long long begin; } e[32];
  int n;
  long long total;
};
__global__ void split_hi_lo_batch_kernel(const SplitBatch b) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < b.total; t += (long long)gridDim.x * blockDim.x) {
    int i = 0;
    while (i + 1 < b.n && t >= b.e[i + 1].begin) ++i;
    const SplitBatch::Entry& en = b.e[i];
    const long long u = t - en.begin, plane = (long long)en.rows_pad * en.K;
    const float v = u / en.K < en.rows ? en.w[u] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    en.out[u] = hi;
    en.out[plane + u] = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
  }
}

// Core: D[pixel, col] = sum_{tap, c} A[n, (oh + kh - pad) * as, (ow + kw - pad) * as, c] * B[col, tap, c]
//   A: [N, AH, AW, Cin] NHWC fp32 read with pixel stride `as` (as > 1 only for 1x1 taps: the TMA view simply
//      has doubled strides, nothing is gathered), B: [ncols, taps*Cin] fp32, output rows enumerate (N, OH, OW)
//   and land at out[((n*out_H + oh*os)*out_W + ow*os)*ldc + col].
// x3: b is the PREPARED matrix [2 * tc_rows_pad(ncols)][taps*Cin] (hi plane, lo plane) and the 3xTF32 kernel runs.
int tc_conv_core(const float* a, int N, int AH, int AW, int Cin, int as, const float* b, int ncols, int KH, int KW,
                 int pad, int OH, int OW, TcParams p, bool x3, cudaStream_t s) {
  const bool flat = KH == 1 && KW == 1 && pad == 0 && as == 1 && p.os == 1 && OH == AH && OW == AW &&
                    p.out_H == OH && p.out_W == OW;
  // 3xTF32 3x3 / pad 1 on small maps whose {tn, th, tw} tiles would be poorly filled (7x7 ROI maps: 98 of 128 rows):
  // the dense pixel axis of the 1x1 case with the taps as row offsets, out-of-map taps zeroed by the operand splitter
  p.flat_h = p.flat_w = 0;
  bool flat_taps = false;
  if (x3 && !flat && KH == 3 && KW == 3 && pad == 1 && as == 1 && p.os == 1 && OH == AH && OW == AW && p.out_H == OH &&
      p.out_W == OW && !p.stem && OH * OW <= 4096) {
    const int tw = pow2_ceil(OW < 16 ? OW : 16), th = pow2_ceil(OH < BM / tw ? OH : BM / tw);
    const double fill = (double)OW / (((OW + tw - 1) / tw) * tw) * (double)OH / (((OH + th - 1) / th) * th);
    static int flat_mode = -1;             // DD_TC_FLAT3X3=0 keeps the {tn, th, tw} tiles (A/B runs)
    if (flat_mode < 0) {
      const char* e = getenv("DD_TC_FLAT3X3");
      flat_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    flat_taps = flat_mode == 1 && fill < 0.85;
  }
  if (flat_taps) { p.flat_h = OH; p.flat_w = OW; }
  if (flat || flat_taps) {                  // pixels are one dense axis, no tile padding at all
    const long long P = (long long)N * OH * OW;
    DD_CHECK_ARG(P < (1ll << 31));
    N = 1; OH = AH = 1; OW = AW = (int)P;
    p.out_H = 1; p.out_W = (int)P;
    p.tw = BM; p.th = 1; p.tn = 1;
  } else {
    p.tw = pow2_ceil(OW < 16 ? OW : 16);
    p.th = pow2_ceil(OH < BM / p.tw ? OH : BM / p.tw);
    p.tn = BM / (p.tw * p.th);
  }
  p.N = N; p.OH = OH; p.OW = OW;
  p.tiles_w = (OW + p.tw - 1) / p.tw;
  p.tiles_h = (OH + p.th - 1) / p.th;
  const int tiles_n = (N + p.tn - 1) / p.tn;
  p.taps = KH * KW; p.KW = KW; p.pad = pad; p.Cin = Cin; p.cblocks = Cin / BKE; p.Cout = ncols;
  p.m_tiles = tiles_n * p.tiles_h * p.tiles_w;
  const int BN = tc_bn_for(ncols, x3);
  p.n_tiles = (ncols + BN - 1) / BN;
  // CTA pairs (256-pixel x BN tiles, half the B traffic per output) whenever there are at least two pixel tiles
  // and the K loop is long enough to pay for the pair's cluster barriers (measured per layer, TF32: +5..9 % from 18
  // K-iterations up, -5..25 % on the 2..16-iteration 1x1 layers); DD_TC_CTA2=2 forces pairs everywhere (tests)
  // The 3xTF32 kernel never pairs its MMAs: its stage chain (TMA -> splitter -> MMA) is latency-bound, and the pair's
  // cross-CTA hops (remote split arrive, cluster-scope wait, multicast commit) lengthen exactly that chain (measured:
  // RPN 3x3 1.49 ms single, 2.28 ms paired).  It shares the B tile by TMA multicast instead (PAIR 2).
  // DD_TC_CTA2: 0 = single CTAs everywhere, 2 = cta_group::2 pairs everywhere, 3 = multicast pairs everywhere (tests).
  const int k_iters_host = p.taps * p.cblocks;
  int pair = 0;
  if (tc_cta2_enabled() && !p.stem && p.m_tiles >= 2) {
    const int forced = tc_cta2_forced();
    if (forced) pair = (x3 && forced == 1) ? 0 : forced;
    else if (!x3 && k_iters_host >= 18) pair = 1;
    else if (x3 && k_iters_host >= 18) pair = 2;
  }
  p.b_lo_row = x3 ? tc_rows_pad(ncols) : 0;
  DD_CHECK_ARG((long long)p.m_tiles * p.n_tiles < (1ll << 31));
  const uintptr_t align_bits = reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.extra) |
                               reinterpret_cast<uintptr_t>(p.mask);
  p.tma_store = (p.ldc % 4 == 0 && (align_bits & 15) == 0) ? 1 : 0;

  CUtensorMap ma, mb, mc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)AW, (cuuint64_t)AH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 4 * as, (cuuint64_t)AW * Cin * 4 * as, (cuuint64_t)AH * AW * Cin * 4};
    if (as > 1) { dims[1] = (cuuint64_t)OW; dims[2] = (cuuint64_t)OH; }   // view of every as-th pixel of [AH, AW]
    cuuint32_t box[4] = {(cuuint32_t)BKE, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    if (encode_map(&ma, a, 4, dims, strides, box)) return -1;
  }
  {
    const cuuint64_t K = (cuuint64_t)KH * KW * Cin;
    cuuint64_t dims[2] = {K, (cuuint64_t)(x3 ? 2 * tc_rows_pad(ncols) : ncols)};
    cuuint64_t strides[1] = {K * 4};
    cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)(pair ? BN / 2 : BN)};    // a CTA of a pair loads half the columns
    if (encode_map(&mb, b, 2, dims, strides, box)) return -1;
  }
  if (p.tma_store) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)p.ldc * 4 * p.os, (cuuint64_t)p.out_W * p.ldc * 4 * p.os,
                             (cuuint64_t)p.out_H * p.out_W * p.ldc * 4};
    // each epilogue warp stores its own 32 tile rows: a [bn x bh x bw] sub-box of the tile
    const int bw = p.tw < 32 ? p.tw : 32;
    const int bh = p.th < 32 / bw ? p.th : 32 / bw;
    const int bn = 32 / (bw * bh);
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    if (encode_map(&mc, p.out, 4, dims, strides, box)) return -1;
  } else {
    mc = ma;
  }
  CUtensorMap me = ma, mm = ma;
  p.prefetch_side = 0;
  p.onepass = 0;
  {
    const int onepass_mode = tc_onepass_enabled() ? 1 : 0;
    static int onepass_k = -1;             // longest K loop (in 32-channel iterations) that takes the one-pass epilogue
    if (onepass_k < 0) {
      const char* e = getenv("DD_TC_ONEPASS_K");
      onepass_k = e != nullptr ? atoi(e) : 8;
    }
    // measured per layer (profiles/r03_layer_bench_x3_onepass.txt): with a residual it pays up to 8 K-iterations, without
    // one up to 4 — beyond that the pipeline stage the side buffers cost (3 instead of 4 at BN = 128) weighs more
    // (64-column tiles keep four stages with the side buffers: up to 8 iterations without a residual as well — the
    // 64 -> 64 and 256 -> 64 convs of res2)
    const int k_limit = (p.extra || BN == 64) ? onepass_k : onepass_k / 2;
    if (onepass_mode == 1 && x3 && p.tma_store && !p.mask && k_iters_host <= k_limit && BN <= 128)
      p.onepass = 1;
  }
  if (p.onepass && p.extra) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)p.ldc * 4 * p.os, (cuuint64_t)p.out_W * p.ldc * 4 * p.os,
                             (cuuint64_t)p.out_H * p.out_W * p.ldc * 4};
    const int bw = p.tw < 32 ? p.tw : 32;
    const int bh = p.th < 32 / bw ? p.th : 32 / bw;
    const int bn = 32 / (bw * bh);
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    if (encode_map(&me, p.extra, 4, dims, strides, box)) return -1;
  } else if (p.tma_store && (p.extra || p.mask)) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)p.ldc * 4 * p.os, (cuuint64_t)p.out_W * p.ldc * 4 * p.os,
                             (cuuint64_t)p.out_H * p.out_W * p.ldc * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    if (p.extra && encode_map(&me, p.extra, 4, dims, strides, box)) return -1;
    if (p.mask && encode_map(&mm, p.mask, 4, dims, strides, box)) return -1;
    // The producer-side L2 prefetch of the side tiles is kept in the kernel but switched off: once the epilogue
    // issues its side reads one chunk ahead, the prefetch only made the residual come from DRAM twice (ncu: 1.6x
    // read traffic on res2 conv3, 0.192 -> 0.167 ms without it) and changes nothing when both inputs are present.
    p.prefetch_side = 0;
  }
  if (BN == 256) return launch_tc<256>(ma, mb, mc, me, mm, p, x3, pair, s);
  if (BN == 128) return launch_tc<128>(ma, mb, mc, me, mm, p, x3, pair, s);
  return launch_tc<64>(ma, mb, mc, me, mm, p, x3, pair, s);
}


template <int BN, bool X3>
int launch_wgrad(const CUtensorMap& mg, const CUtensorMap& mx, const TcWgradParams& p, dim3 grid, cudaStream_t s) {
  using L = WgradSmem<BN, X3>;
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  wgrad_tc_kernel<BN, X3><<<grid, X3 ? NUM_THREADS_X3 : NUM_THREADS, L::kTotal, s>>>(mg, mx, p);
  DD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------- 7x7/2 stem
// The stem conv (3 -> 64, 7x7, stride 2, pad 3; resnet.py:317-336) has Cin = 3, far too narrow for a TMA box.
// Re-cast: copy the NCHW image once into a zero-haloed NHWC4 buffer P[N, H+6, W+8, 4]; then the 7 taps of one
// filter ROW that an output pixel reads are 8 consecutive pixels x 4 channels = 32 contiguous floats of P
// (the 8th pixel and the 4th channel meet zero weights).  A TMA view with OVERLAPPING windows — pixel stride
// 32 B, row pairs split into (row/2, parity) so that the stride-2 row walk is a plain dimension — hands the
// tensor cores a K-major [128 pixels x 32] tile per filter row: an implicit GEMM with K = 7 x 32 = 224.
__global__ void stem_pad_kernel(const float* __restrict__ x, float* __restrict__ P, int N, int H, int W, int hp, int wp) {
  // one thread per padded pixel; reads of the three planes are coalesced along w
  const long long total = (long long)N * hp * wp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(t % wp) - 3;
    const int h = (int)((t / wp) % hp) - 3;
    const int n = (int)(t / ((long long)wp * hp));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w >= 0 && w < W && h >= 0 && h < H) {
      const size_t plane = (size_t)H * W, o = (size_t)n * 3 * plane + (size_t)h * W + w;
      v.x = __ldg(x + o); v.y = __ldg(x + o + plane); v.z = __ldg(x + o + 2 * plane);
    }
    reinterpret_cast<float4*>(P)[t] = v;
  }
}

// W2[co][kh][kw*4 + c] = w[co][kh][kw][c] (OHWI, 7x7x3), zero for kw == 7 or c == 3
__global__ void stem_weight_kernel(const float* __restrict__ w, float* __restrict__ w2, int Cout) {
  const int total = Cout * 7 * 32;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int c = t & 3, kw = (t >> 2) & 7, kh = (t >> 5) % 7, co = t / (7 * 32);
    w2[t] = (c < 3 && kw < 7) ? w[((co * 7 + kh) * 7 + kw) * 3 + c] : 0.f;
  }
}

}  // namespace

extern "C" size_t dd_stem_workspace_bytes(int N, int H, int W, int Cout) {
  (void)Cout;   // padded image + packed weights + their hi / lo planes (64 rows each) for the 3xTF32 mode
  return sizeof(float) * ((size_t)N * (H + 6) * (W + 8) * 4 + 64 + (size_t)3 * 64 * 7 * 32) + 256;
}

extern "C" int dd_stem_conv7x7s2_forward(const float* x_nchw, const float* w_ohwi, const float* scale,
                                         const float* bias, float* y, int N, int H, int W, int Cout, int act,
                                         int impl, void* workspace, void* stream) {
  const bool x3 = impl == DD_IMPL_TCGEN05_X3;
  DD_CHECK_ARG(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && Cout > 0 && Cout % 4 == 0 && Cout <= 64);
  DD_CHECK_ARG(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
               (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  cudaStream_t s = dd::S(stream);
  const int hp = H + 6, wp = W + 8, OH = H / 2, OW = W / 2;
  float* P = (float*)workspace;
  float* w2 = P + (((size_t)N * hp * wp * 4 + 63) / 64) * 64;
  stem_pad_kernel<<<dd::grid_for((long long)N * hp * wp, 256), 256, 0, s>>>(x_nchw, P, N, H, W, hp, wp);
  DD_LAUNCHED();
  stem_weight_kernel<<<(Cout * 7 * 32 + 255) / 256, 256, 0, s>>>(w_ohwi, w2, Cout);
  DD_LAUNCHED();
  const float* bmat = w2;
  if (x3) {
    float* w3 = w2 + 64 * 7 * 32;
    split_hi_lo_kernel<<<dd::grid_for(64ll * 224, 256), 256, 0, s>>>(w2, w3, Cout, 64, 224);
    DD_LAUNCHED();
    bmat = w3;
  }

  TcParams p = {};
  p.out = y; p.scale = scale; p.bias = bias; p.relu = act == DD_ACT_RELU;
  p.N = N; p.OH = OH; p.OW = OW;
  p.tw = 16; p.th = 8; p.tn = 1;
  p.tiles_w = (OW + 15) / 16; p.tiles_h = (OH + 7) / 8;
  p.m_tiles = N * p.tiles_h * p.tiles_w; p.n_tiles = 1;
  p.tma_store = 1; p.stem = 1;
  p.onepass = 0;       // (measured: the one-pass epilogue makes the stem slower, 228 us against 180)
  p.out_H = OH; p.out_W = OW; p.os = 1; p.ldc = Cout; p.Cout = Cout;
  p.taps = 7; p.KW = 1; p.pad = 0; p.cblocks = 1; p.Cin = 32;
  p.b_lo_row = x3 ? 64 : 0;
  CUtensorMap ma, mb, mc;
  {
    // (32 floats of the window, ow, row pair, row parity, n)
    cuuint64_t dims[5] = {32, (cuuint64_t)OW, (cuuint64_t)(hp / 2), 2, (cuuint64_t)N};
    cuuint64_t strides[4] = {32, (cuuint64_t)wp * 32, (cuuint64_t)wp * 16, (cuuint64_t)hp * wp * 16};
    cuuint32_t box[5] = {32, 16, 8, 1, 1};
    if (encode_map(&ma, P, 5, dims, strides, box)) return -1;
  }
  {
    cuuint64_t dims[2] = {224, (cuuint64_t)(x3 ? 128 : Cout)};
    cuuint64_t strides[1] = {224 * 4};
    cuuint32_t box[2] = {32, 64};
    if (encode_map(&mb, bmat, 2, dims, strides, box)) return -1;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)OW * Cout * 4, (cuuint64_t)OH * OW * Cout * 4};
    cuuint32_t box[4] = {32, 16, 2, 1};
    if (encode_map(&mc, y, 4, dims, strides, box)) return -1;
  }
  return launch_tc<64>(ma, mb, mc, ma, ma, p, x3, 0, s);
}

namespace {
}  // namespace

extern "C" int dd_tcgen05_built(void) { return 1; }

// Prepare the dgrad weights W' of n layers (what dd_conv2d_dgrad does per call when prepared == 0) in one launch per
// 16 layers.  ws[i] must hold dd_conv2d_dgrad_workspace_bytes(Cin[i], Cout[i], KH[i], KW[i]) bytes.
int dd_tc_dgrad_prepare_batch(int n, const float* const* w, const float* const* scale, float* const* ws,
                              const int* Cin, const int* Cout, const int* KH, const int* KW, bool x3, cudaStream_t s) {
  for (int base = 0; base < n; base += 16) {
    PrepBatch b = {};
    b.n = n - base < 16 ? n - base : 16;
    int tiles = 0;
    for (int i = 0; i < b.n; ++i) {
      const int j = base + i, taps = KH[j] * KW[j];
      const long long plane = (long long)tc_rows_pad(Cin[j]) * taps * Cout[j];
      if (x3 && tc_rows_pad(Cin[j]) != Cin[j]) DD_CUDA(cudaMemsetAsync(ws[j], 0, sizeof(float) * 2 * plane, s));
      b.e[i].w = w[j]; b.e[i].scale = scale[j]; b.e[i].wt = ws[j];
      b.e[i].Cout = Cout[j]; b.e[i].Cin = Cin[j]; b.e[i].taps = taps; b.e[i].tile_begin = tiles;
      b.e[i].lo_off = x3 ? plane : 0;
      tiles += ((Cin[j] + 31) / 32) * ((Cout[j] + 31) / 32) * taps;
    }
    weight_flip_transpose_batch_kernel<<<tiles, dim3(32, 8), 0, s>>>(b);
    DD_LAUNCHED();
  }
  return 0;
}
int tc_rows_pad_public(int ncols) { return ncols > 64 ? (ncols + 127) / 128 * 128 : 64; }

// mode: 0 forward, 1 dgrad, 2 wgrad
bool dd_tc_supports(int mode, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  (void)N; (void)H; (void)W;
  if (mode == 0) {
    if (Cin % 32 != 0) return false;                                  // stem (Cin = 3) stays on the SIMT arm
    if (stride != 1 && !(KH == 1 && KW == 1 && pad == 0)) return false;
    return true;
  }
  if (mode == 1) {
    if (Cout % 32 != 0 || Cin % 4 != 0) return false;                 // K of the dgrad GEMM is Cout
    if (stride != 1 && !(KH == 1 && KW == 1 && pad == 0)) return false;
    if (stride == 1 && pad != (KH - 1) / 2) return false;             // "same" convs only (all of ResNet/RPN)
    return true;
  }
  // wgrad: both channel counts are TMA inner dimensions
  if (Cin % 32 != 0 || Cout % 32 != 0) return false;
  if (stride != 1 && !(KH == 1 && KW == 1 && pad == 0)) return false;
  return true;
}

// Split the forward weights of n layers into the hi / lo planes the 3xTF32 kernel consumes (what dd_conv2d_forward does
// per call) in one launch per 32 layers.  ws[i] must hold dd_conv2d_forward_workspace_bytes(...) bytes.
int dd_tc_forward_prepare_batch(int n, const float* const* w, float* const* ws, const int* Cin, const int* Cout,
                                const int* KH, const int* KW, cudaStream_t s) {
  for (int base = 0; base < n; base += 32) {
    SplitBatch b = {};
    b.n = n - base < 32 ? n - base : 32;
    long long total = 0;
    for (int i = 0; i < b.n; ++i) {
      const int j = base + i;
      b.e[i].w = w[j]; b.e[i].out = ws[j]; b.e[i].rows = Cout[j]; b.e[i].rows_pad = tc_rows_pad(Cout[j]);
      b.e[i].K = (long long)KH[j] * KW[j] * Cin[j]; b.e[i].begin = total;
      total += (long long)b.e[i].rows_pad * b.e[i].K;
    }
    b.total = total;
    split_hi_lo_batch_kernel<<<dd::grid_for(total, 256), 256, 0, s>>>(b);
    DD_LAUNCHED();
  }
  return 0;
}

int dd_tc_conv2d_forward(const float* x, const float* w, const float* scale, const float* bias, const float* residual,
                         float* y, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                         int act, bool x3, float* ws, bool prepared, cudaStream_t s) {
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  TcParams p = {};
  p.out = y; p.scale = scale; p.bias = bias; p.extra = residual; p.mask = nullptr; p.relu = act == DD_ACT_RELU;
  p.out_H = OH; p.out_W = OW; p.os = 1; p.ldc = Cout;
  const float* b = w;
  if (x3) {                                  // split the weights into TF32-exact hi / lo planes
    DD_CHECK_ARG(ws != nullptr);
    const long long K = (long long)KH * KW * Cin;
    const int rp = tc_rows_pad(Cout);
    if (!prepared) {
      split_hi_lo_kernel<<<dd::grid_for((long long)rp * K, 256), 256, 0, s>>>(w, ws, Cout, rp, K);
      DD_LAUNCHED();
    }
    b = ws;
  }
  // strided 1x1: the TMA view of x addresses every stride-th pixel directly
  return tc_conv_core(x, N, H, W, Cin, stride, b, Cout, KH, KW, pad, OH, OW, p, x3, s);
}

int dd_tc_conv2d_dgrad(const float* gy, const float* w, const float* scale, const float* addend, const float* mask_act,
                       float* gx, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                       float* wt, int prepared, bool x3, cudaStream_t s) {
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  if (!prepared) {
    const long long plane = (long long)tc_rows_pad(Cin) * KH * KW * Cout;
    if (x3 && tc_rows_pad(Cin) != Cin) DD_CUDA(cudaMemsetAsync(wt, 0, sizeof(float) * 2 * plane, s));   // pad rows
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, KH * KW);
    weight_flip_transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(w, scale, wt, Cout, Cin, KH * KW, x3 ? plane : 0);
    DD_LAUNCHED();
  }
  TcParams p = {};
  p.out = gx; p.scale = nullptr; p.bias = nullptr; p.extra = addend; p.mask = mask_act; p.relu = 0;
  p.out_H = H; p.out_W = W; p.ldc = Cin;
  int rc;
  if (stride == 1) {
    p.os = 1;
    // gx[n,h,w,ci] = sum_{kh',kw',co} gy[n, h + kh' - pad', w + kw' - pad', co] * W'[ci, kh', kw', co], pad' = KH-1-pad
    rc = tc_conv_core(gy, N, OH, OW, Cout, 1, wt, Cin, KH, KW, KH - 1 - pad, H, W, p, x3, s);
  } else {
    // 1x1 stride s: rows enumerate gy's pixels, each lands on (oh*s, ow*s); everything else is addend/0
    const long long n = (long long)N * H * W * Cin;
    fill_kernel<<<dd::grid_for(n, 256), 256, 0, s>>>(addend, mask_act, gx, n);
    DD_LAUNCHED();
    p.os = stride;
    rc = tc_conv_core(gy, N, OH, OW, Cout, 1, wt, Cin, 1, 1, 0, OH, OW, p, x3, s);
  }
  return rc;
}

int dd_tc_conv2d_wgrad(const float* gy, const float* x, const float* scale, float* gw, int N, int H, int W, int Cin,
                       int Cout, int KH, int KW, int stride, int pad, int accumulate, void* workspace, bool x3,
                       cudaStream_t s) {
  (void)workspace;
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  DD_CHECK_ARG((reinterpret_cast<uintptr_t>(gw) & 15) == 0);
  if (!accumulate) DD_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Cout * KH * KW * Cin, s));
  TcWgradParams p = {};
  p.gw = gw; p.scale = scale;
  p.Cout = Cout; p.Cin = Cin; p.taps = KH * KW; p.KW = KW; p.pad = pad;
  // pixel axes as the kernel sees them (gy side: PN x PH x PW; x side may be a strided view of the same grid)
  int PN = N, PH = OH, PW = OW;
  cuuint64_t xs_w = (cuuint64_t)Cin * 4 * stride, xs_h = (cuuint64_t)W * Cin * 4 * stride, xs_n = (cuuint64_t)H * W * Cin * 4;
  cuuint64_t x_w = stride == 1 ? W : OW, x_h = stride == 1 ? H : OH;
  if (KH == 1 && KW == 1 && pad == 0 && stride == 1) {        // flat: one dense pixel axis
    const long long P = (long long)N * OH * OW;
    DD_CHECK_ARG(P < (1ll << 31));
    PN = 1; PH = 1; PW = (int)P;
    x_w = (cuuint64_t)P; x_h = 1;
    xs_h = xs_n = (cuuint64_t)P * Cin * 4;
    p.tw = WG_KPIX; p.th = 1; p.tn = 1;
  } else {
    p.tw = pow2_ceil(PW < 16 ? PW : 16);
    p.th = pow2_ceil(PH < WG_KPIX / p.tw ? PH : WG_KPIX / p.tw);
    p.tn = WG_KPIX / (p.tw * p.th);
  }
  p.tiles_w = (PW + p.tw - 1) / p.tw;
  p.tiles_h = (PH + p.th - 1) / p.th;
  p.pix_blocks = ((PN + p.tn - 1) / p.tn) * p.tiles_h * p.tiles_w;
  const int BN = (Cin % 256 == 0 && !x3) ? 256 : (Cin % 128 == 0 ? 128 : (Cin % 64 == 0 ? 64 : 32));
  p.ci_tiles = (Cin + BN - 1) / BN;
  const int co_tiles = (Cout + BM - 1) / BM;
  // split the pixel range so that the persistent CTAs see whole rounds of work items (items = tiles x splits as
  // close as possible to a multiple of the SM count, 2-4 rounds) and every item still runs >= 8 K-iterations
  const int tiles = co_tiles * p.ci_tiles * p.taps;
  const int sms = dd::sm_budget();
  int splits = 1;
  {
    const int max_splits = (p.pix_blocks + 7) / 8;
    double best = -1.0;
    for (int cand = 1; cand <= max_splits && (long long)cand * tiles <= 6ll * sms; ++cand) {
      const long long items = (long long)cand * tiles;
      const long long rounds = (items + sms - 1) / sms;
      double eff = (double)items / (double)(rounds * sms);
      if (rounds < 2) eff *= 0.9;                       // one round: no overlap of reductions with the next item
      if (eff > best + 1e-9) { best = eff; splits = cand; }
    }
  }
  // 3xTF32: the tensor core's accumulator rounds toward zero at every step, so a CTA accumulates at most 64 pixel
  // blocks (768 steps, ~2e-5 bias) before its partial sum joins the others through round-to-nearest adds in L2
  if (x3 && splits < (p.pix_blocks + 63) / 64) splits = (p.pix_blocks + 63) / 64;
  if (splits < 1) splits = 1;
  p.blocks_per_split = (p.pix_blocks + splits - 1) / splits;
  splits = (p.pix_blocks + p.blocks_per_split - 1) / p.blocks_per_split;

  CUtensorMap mg, mx;
  {
    cuuint64_t dims[5] = {32, (cuuint64_t)PW, (cuuint64_t)PH, (cuuint64_t)PN, (cuuint64_t)(Cout / 32)};
    cuuint64_t strides[4] = {(cuuint64_t)Cout * 4, (cuuint64_t)PW * Cout * 4, (cuuint64_t)PH * PW * Cout * 4, 128};
    cuuint32_t box[5] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn, 4};
    if (encode_map(&mg, gy, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
  }
  {
    // strided 1x1: a view of every stride-th pixel of x (no gather pass)
    cuuint64_t dims[5] = {32, x_w, x_h, (cuuint64_t)PN, (cuuint64_t)(Cin / 32)};
    cuuint64_t strides[4] = {xs_w, xs_h, xs_n, 128};
    cuuint32_t box[5] = {32, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn, (cuuint32_t)(BN / 32)};
    if (encode_map(&mx, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
  }
  p.co_tiles = co_tiles; p.splits = splits;
  const long long items = (long long)p.ci_tiles * p.taps * co_tiles * splits;
  DD_CHECK_ARG(items < (1ll << 31));
  dim3 grid((unsigned)(items < sms ? items : sms));
  int rc;
  if (x3) {
    if (BN == 128) rc = launch_wgrad<128, true>(mg, mx, p, grid, s);
    else if (BN == 64) rc = launch_wgrad<64, true>(mg, mx, p, grid, s);
    else rc = launch_wgrad<32, true>(mg, mx, p, grid, s);
  } else if (BN == 256) rc = launch_wgrad<256, false>(mg, mx, p, grid, s);
  else if (BN == 128) rc = launch_wgrad<128, false>(mg, mx, p, grid, s);
  else if (BN == 64) rc = launch_wgrad<64, false>(mg, mx, p, grid, s);
  else rc = launch_wgrad<32, false>(mg, mx, p, grid, s);
  return rc;
}
