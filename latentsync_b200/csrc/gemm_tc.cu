// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   warps 0..7  : epilogue       (tcgen05.ld -> bias / time-embedding / residual / GEGLU / SiLU -> 64B-swizzled smem
//                                 slab -> TMA bulk store; two groups of 4 warps take alternate 32-column slabs)
//   warp 8      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx), one lane
//   warp 9      : TMEM allocator + tcgen05.mma issuer (fp16 x fp16 -> fp32 in TMEM), one lane
//
// Two TMEM accumulators (double buffered) let the epilogue of tile i overlap the main loop of tile i+1.
//
// What bounds the main loop (measured with in-kernel clock64 probes, profiles/r1_gemm_issue_loop.txt): not HBM, L2 or
// shared memory but the two single-thread loops - one mbarrier round trip plus four MMA issues per 64-wide k-block
// cost ~650 cycles in the first version against 512 cycles of tensor work at N = 256.  Hence:
//   * both loops run on ONE lane of the two HIGHEST warp ids (the sub-partition arbiter prefers high warp ids and the
//     epilogue warps poll with nanosleep back-off, so they cannot steal issue slots),
//   * each iteration is one asm block that first probes the NEXT stage's mbarrier (try_wait) and only then issues the
//     TMA loads / MMAs + commit of the current stage, so the probe's latency hides behind the issue work.
//
// CTA pairs (cta_group::2): a cluster of two CTAs computes a 256 x N tile; each CTA stages its own 128 rows of A and
// HALF of the B rows, the leader's tcgen05.mma.cta_group::2 reads both halves.  Halves the B traffic per MMA.  The
// tile width N (any multiple of 32 up to 256) and the pairing are chosen per launch by a cost model on the host.
//
// The A operand of a 3x3 convolution is never materialised: for filter tap (dy, dx) the producer issues a 4-D
// TMA box load of the channels-last activation shifted by (dy, dx); out-of-bounds coordinates are zero-filled by
// the TMA unit, which is exactly the conv's zero padding.  Extra K segments implement the ResnetBlock3D 1x1
// shortcut and the skip-concat without copies.
//
// Replaces (reference, all ATen/cuDNN/cuBLAS library calls): InflatedConv3d.forward latentsync/models/resnet.py:10-18,
// nn.Linear in attention.py:230-235 / motion_module.py:102,124, Conv2d proj_in/out attention.py:55,80,
// conv_shortcut resnet.py:180,219-221, diffusers FeedForward/GEGLU (attention.py:171).
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>
#include <stdlib.h>

namespace ls {

extern std::atomic<int64_t> g_launch_count;

// Which lane runs the two control loops (TMA producer, MMA issuer).  Under `if (lane == 0)` ptxas cannot tell that a
// single lane is active: it keeps the operands in vector registers, moves them to uniform registers (R2UR) every
// k-block and wraps each UTMALDG / UTCHMMA / UTCBAR in an ELECT + BRA.U.ANY waterfall loop (~75 / ~95 dependent
// instructions per k-block).  With the WHOLE loop under one `elect.sync` predicate it knows, and emits the uniform
// instructions back to back (~45 per k-block).  -DLS_GEMM_LANE0 keeps the old form for A/B runs.
#ifndef LS_GEMM_DEFAULT_BRES
#define LS_GEMM_DEFAULT_BRES 0
#endif
#ifndef LS_GEMM_DEFAULT_KBS
#define LS_GEMM_DEFAULT_KBS 1
#endif
#ifdef LS_GEMM_LANE0
#define LS_ONE_LANE() (lane == 0)
#else
#define LS_ONE_LANE() elect_one()
#endif

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements: one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 320;
constexpr int MAX_STAGES = 8;
constexpr int PRODUCER_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int WSLAB_BYTES = 32 * 64;           // one warp's share of a slab: 32 rows x 32 fp16 columns, SWIZZLE_64B
constexpr int STAGING_BYTES = 8 * 2 * WSLAB_BYTES;  // 8 epilogue warps x (32 rows x 128 bytes): a pair of slabs each
constexpr int BIAS_BYTES = 2 * 256 * 4;        // one bias row of 256 floats per epilogue group
constexpr int COLSUM_BYTES = 2 * 256 * 4;      // folded-LayerNorm launches: one col_sum row per epilogue group
constexpr int GN_SM_BYTES_PER_COL = 2 * 4 * 8;  // LN == 3: float2 column sums [2 slots][4 row quarters][BN]
constexpr int ACC_COLS = 256;                  // TMEM columns per accumulator
constexpr int TMEM_COLS = 512;                 // two accumulators: the whole TMEM (one CTA per SM anyway)
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address (-> pair leader)

// n / d for 0 <= n < 2^31 without the ~40-instruction integer-division sequence (the epilogue warps decode a tile per
// loop trip; the probes showed ~1800 clocks between the last slab of one tile and the first instruction after the
// decode of the next).  Magic numbers by the round-up method: q = umulhi(n, mul) >> shr, d == 1 handled apart.
struct FastDiv {
  uint32_t mul, shr, d;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) {
    f.mul = 0;
    f.shr = 0;
    f.d = 1;
    return f;
  }
  uint32_t lg = 0;
  while ((1ull << lg) < d) ++lg;
  const uint32_t pw = 31 + lg;
  f.mul = (uint32_t)(((1ull << pw) + d - 1) / d);
  f.shr = pw - 32;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

struct GemmKParams {
  CUtensorMap mapA[LS_GEMM_MAX_SEG];
  CUtensorMap mapB;
  CUtensorMap mapOut;    // fp16 [M][N_out] output, box 64 columns x 32 rows (one warp's slab pair), SWIZZLE_128B
  CUtensorMap mapOut32;  // same tensor, box 32 columns x 32 rows (trailing single slab of an odd count), SWIZZLE_64B
  int nseg;
  int seg_taps[LS_GEMM_MAX_SEG];
  int seg_cblk[LS_GEMM_MAX_SEG];
  int bw, bh, bn;  // TMA box in pixels: bw * bh * bn == 128
  int H, W, nimg;
  int tiles_x, tiles_y;
  FastDiv fd_tiles_x, fd_tiles_y, fd_n_tiles, fd_splits, fd_bias_div;
  int m_tiles, n_tiles, num_kb;
  int N;
  int BN;  // tile width (multiple of 32, <= 256)
  int b_batched;
  int stages;
  int kbs;  // 64-wide k-blocks per pipeline stage (1 or 2)
  int bres;  // 1: the CTA's B tile (all k-blocks) stays resident in shared memory across its M tiles; stages hold A only
  int slab_single;  // 1: every 32-column slab is staged and stored on its own (2 KB per warp: 16 KB of staging instead of 32)
  const float* bias;
  int bias_div;
  int bias_ld;
  const __half* residual;
  int ldr;
  void* out;
  int ldo;
  int flags;
  int tma_store;
  int epi_alt;           // 1: the two epilogue groups take alternate tiles (launches with >= 3 tiles per CTA)
  int splits;            // split-K factor: `splits` CTAs share one output tile, each reducing `kb_per` k-blocks
  int kb_per;
  float* ws;             // split-K partials, fp32 [splits][M][N] (library-owned scratch)
  unsigned int* tickets; // one arrival counter per output tile (self-resetting)
  int64_t M;  // total output rows
  // LayerNorm folded into this GEMM (LsGemmArgs.col_sum / row_partials_in): out = rstd (x W'^T) - rstd mean col_sum + bias,
  // (mean, rstd) of row m from the producer's partials ln_in[i * ln_in_stride + m], i < ln_nparts_in, summed in order
  const float2* ln_in;
  const float* colsum;
  int ln_nparts_in;
  int64_t ln_in_stride;
  float ln_eps, ln_inv_k;
  // ... and the producer side (LsGemmArgs.row_partials_out): part 3 nt + k of row m = (sum, sum of squares) of the fp16
  // values stored into slab range k of n-tile nt
  float2* ln_out;
  int64_t ln_out_stride;
  // GroupNorm statistics of the OUTPUT (LsGemmArgs.gn_partials_out, LN == 3): per 128-row tile and per `gn_unit` consecutive
  // output columns the (sum, sum of squares) of the fp16 values stored, gn_out[(m0 / 128) * gn_ld + column / gn_unit]
  float2* gn_out;
  int gn_unit, gn_ld, gn_upt;  // gn_upt = BN / gn_unit: units per tile
  // Sub-pixel phase of "nearest x2 upsample -> 3x3 convolution" (LsGemmArgs.up2): a segment with 4 taps reads the 2 x 2
  // window of the LOW-resolution image that starts at (tap_dy0, tap_dx0) in {-1, 0}^2; the tile's rows are low-resolution
  // pixels and leave through 4-D output maps over [column][x][y][image] of the high-resolution tensor with strides of two
  // pixels / two rows, based at the phase's pixel (py, px).  up2_bx x up2_by = 32 pixels: the box of one warp's rows.
  int up2, tap_dy0, tap_dx0, up2_bx, up2_by;
  // Stride-2 3x3 convolution read in place (LsGemmArgs.stride2): tile coordinates are OUTPUT pixels, the A maps cover the
  // INPUT image with element strides of two pixels / rows, so box coordinates are input pixels: cs * x0 + dx with cs = 2;
  // tap9_off = -pad (first tap offset of a 9-tap segment: -1 for zero padding 1, 0 for the (0, 1, 0, 1) padding)
  int cs, tap9_off;
};

// ----------------------------------------------------------------------------------------- pair-mode PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CTAS>
__device__ __forceinline__ void tmem_alloc_g(uint32_t* dst_smem, uint32_t ncols) {
  if constexpr (CTAS == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr, uint32_t ncols) {
  if constexpr (CTAS == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this smem offset in every CTA of the pair
template <int CTAS>
__device__ __forceinline__ void umma_commit_g(uint64_t* bar) {
  if constexpr (CTAS == 1) {
    umma_commit(bar);
  } else {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
  }
}
// arrive on the pair leader's copy of `bar` (for the leader itself this is its own barrier)
template <int CTAS>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  if constexpr (CTAS == 1) {
    mbar_arrive(bar);
  } else {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK)
                 : "memory");
  }
}

// One k-block of the MMA issuer as ONE asm block: probe the next stage's `full` barrier, issue the four K=16 MMAs of
// this stage (smem descriptors advance by 32 bytes = +2 in the 16-byte-unit address field), commit to this stage's
// `empty` barrier, and only then consume the probe result.  Returns 1 if the next stage has already landed.
template <int CTAS>
__device__ __forceinline__ uint32_t mma_kblock(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate, uint32_t empty_bar, uint32_t next_full_bar,
                                               uint32_t next_parity) {
  uint32_t ready;
  if constexpr (CTAS == 1) {
    asm volatile(
        "{\n"
        ".reg .pred P, ACC, T;\n"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%7], %8;\n"
        "setp.ne.b32 ACC, %5, 0;\n"
        "setp.eq.b32 T, %4, %4;\n"
        "add.s64 a1, %2, 2;\n add.s64 b1, %3, 2;\n"
        "add.s64 a2, %2, 4;\n add.s64 b2, %3, 4;\n"
        "add.s64 a3, %2, 6;\n add.s64 b3, %3, 6;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, ACC;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a1, b1, %4, T;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a2, b2, %4, T;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a3, b3, %4, T;\n"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}"
        : "=r"(ready)
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(empty_bar), "r"(next_full_bar),
          "r"(next_parity)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred P, ACC, T;\n"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n"
        ".reg .b16 m;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%7], %8;\n"
        "setp.ne.b32 ACC, %5, 0;\n"
        "setp.eq.b32 T, %4, %4;\n"
        "mov.b16 m, 3;\n"
        "add.s64 a1, %2, 2;\n add.s64 b1, %3, 2;\n"
        "add.s64 a2, %2, 4;\n add.s64 b2, %3, 4;\n"
        "add.s64 a3, %2, 6;\n add.s64 b3, %3, 6;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %4, ACC;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a1, b1, %4, T;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a2, b2, %4, T;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a3, b3, %4, T;\n"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], m;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}"
        : "=r"(ready)
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(empty_bar), "r"(next_full_bar),
          "r"(next_parity)
        : "memory");
  }
  return ready;
}

// One k-block of the TMA producer as ONE asm block: probe the next stage's `empty` barrier, arm this stage's `full`
// barrier (pair leader only) and issue the A (4-D box) and B (3-D box) loads.  In pair mode the data lands in THIS
// CTA's smem while the transaction bytes are counted on the LEADER's barrier (address with the CTA-rank bit cleared).
template <int CTAS>
__device__ __forceinline__ uint32_t produce_kblock(uint32_t sa, uint32_t sb, const CUtensorMap* mapA,
                                                   const CUtensorMap* mapB, uint32_t full_bar, uint32_t tx_bytes,
                                                   int a0, int a1, int a2, int a3, int b0, int b1, int b2,
                                                   uint32_t next_empty_bar, uint32_t next_parity) {
  uint32_t ready;
  if constexpr (CTAS == 1) {
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%14], %15;\n"
        "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%5], %6;\n"
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%3, {%7, %8, %9, %10}], [%5];\n"
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%4, {%11, %12, %13}], [%5];\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}"
        : "=r"(ready)
        : "r"(sa), "r"(sb), "l"(reinterpret_cast<uint64_t>(mapA)), "l"(reinterpret_cast<uint64_t>(mapB)),
          "r"(full_bar), "r"(tx_bytes), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2),
          "r"(next_empty_bar), "r"(next_parity)
        : "memory");
  } else {
    // tx_bytes == 0 on the non-leader CTA: it only issues its loads
    asm volatile(
        "{\n"
        ".reg .pred P, L;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%14], %15;\n"
        "setp.ne.b32 L, %6, 0;\n"
        "@L mbarrier.arrive.expect_tx.shared::cta.b64 _, [%5], %6;\n"
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%3, {%7, %8, %9, %10}], [%16];\n"
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%4, {%11, %12, %13}], [%16];\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}"
        : "=r"(ready)
        : "r"(sa), "r"(sb), "l"(reinterpret_cast<uint64_t>(mapA)), "l"(reinterpret_cast<uint64_t>(mapB)),
          "r"(full_bar), "r"(tx_bytes), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2),
          "r"(next_empty_bar), "r"(next_parity), "r"(full_bar & PEER_MASK)
        : "memory");
  }
  return ready;
}

// ---- two k-blocks per pipeline stage (p.kbs == 2): the pieces of produce_kblock / mma_kblock as separate calls.
// One mbarrier round trip (probe + arm + commit) per stage costs each single-lane loop ~370 clocks whatever it issues
// (profiles/r2_gemm_ablate_elect.txt, "nothing": the latency of two dependent mbarrier operations, not instruction
// count); with 128 K columns per stage that cost is paid once per two swizzle atoms.
__device__ __forceinline__ uint32_t mbar_probe(uint32_t bar, uint32_t parity) {
  uint32_t ready;
  asm volatile(
      "{\n.reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n}"
      : "=r"(ready) : "r"(bar), "r"(parity) : "memory");
  return ready;
}
template <int CTAS>
__device__ __forceinline__ void tma_a_4d(uint32_t dst, const CUtensorMap* map, uint32_t full_bar, int c0, int c1, int c2,
                                         int c3) {
  if constexpr (CTAS == 1)
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(full_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(full_bar & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tma_b_3d(uint32_t dst, const CUtensorMap* map, uint32_t full_bar, int c0, int c1, int c2) {
  if constexpr (CTAS == 1)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(full_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(full_bar & PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// the four K = 16 MMAs of one 64-wide k-block (no probe, no commit)
template <int CTAS>
__device__ __forceinline__ void mma_quad(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (CTAS == 1)
    asm volatile(
        "{\n.reg .pred ACC, T;\n.reg .b64 a1, a2, a3, b1, b2, b3;\n"
        "setp.ne.b32 ACC, %4, 0;\n setp.eq.b32 T, %3, %3;\n"
        "add.s64 a1, %1, 2;\n add.s64 b1, %2, 2;\n add.s64 a2, %1, 4;\n add.s64 b2, %2, 4;\n"
        "add.s64 a3, %1, 6;\n add.s64 b3, %2, 6;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, ACC;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, T;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, T;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, T;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n.reg .pred ACC, T;\n.reg .b64 a1, a2, a3, b1, b2, b3;\n"
        "setp.ne.b32 ACC, %4, 0;\n setp.eq.b32 T, %3, %3;\n"
        "add.s64 a1, %1, 2;\n add.s64 b1, %2, 2;\n add.s64 a2, %1, 4;\n add.s64 b2, %2, 4;\n"
        "add.s64 a3, %1, 6;\n add.s64 b3, %2, 6;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, ACC;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, T;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, T;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, T;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void decode_m_tile(const GemmKParams& p, int mt, int& x0, int& y0, int& i0) {
  const int rest = (int)fdiv((uint32_t)mt, p.fd_tiles_x);
  const int tx = mt - rest * p.tiles_x;
  const int tn = (int)fdiv((uint32_t)rest, p.fd_tiles_y);
  const int ty = rest - tn * p.tiles_y;
  x0 = tx * p.bw;
  y0 = ty * p.bh;
  i0 = tn * p.bn;
}

// ------------------------------------------------------------------------- legacy (direct global store) epilogue
// bias / residual / activation / store for 32 consecutive output columns of one row
__device__ __forceinline__ void epilogue_store32(const GemmKParams& p, int64_t m, int n_base, int n_total,
                                                 float (&f)[32]) {
  const int nvalid = min(32, n_total - n_base);
  const bool full = (nvalid == 32);
  if (p.residual != nullptr) {
    const __half* rr = p.residual + m * (int64_t)p.ldr + n_base;
    if (full && (p.ldr & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = *reinterpret_cast<const uint4*>(rr + j * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __half22float2(h2[e]);
          f[j * 8 + e * 2] += t.x;
          f[j * 8 + e * 2 + 1] += t.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) f[j] += __half2float(rr[j]);
    }
  }
  if (p.flags & LS_EPI_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
  }
  if (p.flags & LS_EPI_OUT_F32) {
    float* o = reinterpret_cast<float*>(p.out) + m * (int64_t)p.ldo + n_base;
    if (full && (p.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(o + j * 4) = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = f[j];
    }
  } else {
    __half* o = reinterpret_cast<__half*>(p.out) + m * (int64_t)p.ldo + n_base;
    if (full && (p.ldo & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
        *reinterpret_cast<uint4*>(o + j * 8) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j] = __float2half_rn(f[j]);
    }
  }
}

__device__ __forceinline__ void add_bias32(const GemmKParams& p, int64_t m, int n_base, float (&f)[32]) {
  if (p.bias == nullptr) return;
  const float* b = p.bias + (p.bias_div > 0 ? (m / p.bias_div) * (int64_t)p.bias_ld : 0) + n_base;
  const int nvalid = min(32, p.N - n_base);
  if (nvalid == 32 && (p.bias_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(b + j * 4));
      f[j * 4] += t.x;
      f[j * 4 + 1] += t.y;
      f[j * 4 + 2] += t.z;
      f[j * 4 + 3] += t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) f[j] += __ldg(b + j);
  }
}


// ------------------------------------------------------------------------- LN == 3: column sums of a staged slab
// (sum, sum of squares) over the warp's 32 rows of the fp16 values just staged for the TMA store - exactly the values the
// consuming GroupNorm will read.  The association is the same for both staging layouts, so that the partials do not depend
// on how a launch happens to split its slabs: rows 0..15 ascending, rows 16 + (i ^ 1) for i = 0..15, then first + second.
// Pair layout (128-byte rows, SWIZZLE_128B): lane l owns columns 2l, 2l + 1 of the 64; every load touches one row = 32 banks.
__device__ __forceinline__ void gn_colsum_pair(const uint8_t* st, int lane, float2* dst) {
  // packed fp32 pairs (add.f32x2 / fma.rn.f32x2 of sm_100: the same roundings as two scalar operations, half the issue slots)
  float2 sa = make_float2(0.f, 0.f), qa = sa, sb = sa, qb = sa;
  const int ch = lane >> 2, sub = (lane & 3) * 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int ra = i, rb = 16 + (i ^ 1);
    const __half2 ha = *reinterpret_cast<const __half2*>(st + ra * 128 + ((ch ^ (ra & 7)) << 4) + sub);
    const __half2 hb = *reinterpret_cast<const __half2*>(st + rb * 128 + ((ch ^ (rb & 7)) << 4) + sub);
    const float2 a = __half22float2(ha), b = __half22float2(hb);
    sa = __fadd2_rn(sa, a);
    qa = __ffma2_rn(a, a, qa);
    sb = __fadd2_rn(sb, b);
    qb = __ffma2_rn(b, b, qb);
  }
  const float2 s = __fadd2_rn(sa, sb), q = __fadd2_rn(qa, qb);
  *reinterpret_cast<float4*>(dst + 2 * lane) = make_float4(s.x, q.x, s.y, q.y);
}
// Single-slab layout (64-byte rows, SWIZZLE_64B): lane l owns columns 2 (l & 15), + 1 over the row half l >> 4; the two
// halves read rows of opposite parity (different 16-bank halves), then meet by one shuffle.
__device__ __forceinline__ void gn_colsum_single(const uint8_t* st, int lane, float2* dst) {
  float2 s = make_float2(0.f, 0.f), q = s;
  const int cp = lane & 15, hi = lane >> 4;
  const int ch = cp >> 2, sub = (cp & 3) * 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = hi ? 16 + (i ^ 1) : i;
    const __half2 h = *reinterpret_cast<const __half2*>(st + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4) + sub);
    const float2 a = __half22float2(h);
    s = __fadd2_rn(s, a);
    q = __ffma2_rn(a, a, q);
  }
  const float t0 = __shfl_xor_sync(0xffffffffu, s.x, 16), t1 = __shfl_xor_sync(0xffffffffu, s.y, 16);
  const float u0 = __shfl_xor_sync(0xffffffffu, q.x, 16), u1 = __shfl_xor_sync(0xffffffffu, q.y, 16);
  // (first half + second half, whichever lane computes it: the same bits as gn_colsum_pair)
  if (hi == 0) *reinterpret_cast<float4*>(dst + 2 * cp) = make_float4(s.x + t0, q.x + u0, s.y + t1, q.y + u1);
}

// ----------------------------------------------------------------------------------------------------- kernel body
// GEGLU is a template parameter: the plain instantiation carries no gate accumulator (161 instead of 168 registers - 168
// is the limit with 10 warps, 3 on one SM sub-partition).  Tried on top of it and dropped: keeping the NEXT slab's
// tcgen05.ld in flight while the current one is converted (needs ~190 registers: spills, 71 vs 50 us on the K = 64,
// N = 2560 probe); setmaxnreg with a third, shrunken warpgroup for the producer / issuer (ptxas then spills in the
// producer / issuer loops at every split tried: 224/64, 208/96, 200/112, 192/128).
// LN selects the folded-LayerNorm epilogues (separate instantiations: the plain kernels do not pay for them - the first
// version of this fold, round 1, put the code into every launch and lost more there than the LayerNorm kernels cost):
//   0 plain;  1 consumer: per-row scale / shift from the producer's partials;  2 producer: emits the partials;
//   3 GroupNorm producer: per 128-row tile and column unit the (sum, sum of squares) of the stored values (gn_out).
template <int CTAS, bool GEGLU, int LN>
__device__ __forceinline__ void gemm_body(const GemmKParams& p) {
  const int BN = p.BN;
  const int b_rows = BN / CTAS;  // B rows staged by each CTA
  const int kblock_bytes = A_STAGE_BYTES + b_rows * 128;  // one 64-wide k-block: A atom + B atom
  const int stage_bytes = p.kbs * kblock_bytes;           // kbs == 2: [A0][A1][B0][B1]
  // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 for a CTA pair)
  const uint32_t idesc = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t((BM * CTAS) >> 4) << 24);

  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  const int stages = p.stages;
  // weight-resident mode: [B: num_kb x (b_rows x 128 B)] [A ring: stages x 16 KB] [staging] ...; otherwise [ring of A+B stages]
  const int bres_bytes = p.bres ? p.num_kb * b_rows * 128 : 0;
  const int ring_stage_bytes = p.bres ? A_STAGE_BYTES : stage_bytes;
  uint8_t* staging = smem + bres_bytes + stages * ring_stage_bytes;  // 1024-byte aligned
  const int staging_bytes = p.slab_single ? STAGING_BYTES / 2 : STAGING_BYTES;
  float* bias_sm = reinterpret_cast<float*>(staging + staging_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + staging_bytes + BIAS_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* b_full = tmem_empty + 3;  // weight-resident mode: the B tile has landed (8 bytes after the TMEM slot)
  float* colsum_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 192);  // LN == 1 launches only
  float2* gn_sm = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(full_bar) + 192);     // LN == 3 launches only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (CTAS == 2) ? (int)cluster_ctarank() : 0;
  const bool leader = (rank == 0);
  const int unit = blockIdx.x / CTAS;  // CTA (or CTA pair) index
  const int nunits = gridDim.x / CTAS;
  const int m_units = (p.m_tiles + CTAS - 1) / CTAS;
  const int total_tiles = m_units * p.n_tiles;
  const int total_work = total_tiles * p.splits;  // work item w = (tile w / splits, K-range w % splits)

  if (warp == PRODUCER_WARP && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapB);
    if (p.tma_store) {
      tma_prefetch_desc(&p.mapOut);
      tma_prefetch_desc(&p.mapOut32);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      // one arrival per epilogue warp that reads the accumulator: all 8 (split-K, legacy path) or the 4 of the group
      // that owns the tile (epi_alt: the two groups take alternate tiles)
      mbar_init(&tmem_empty[a], (p.epi_alt ? 4 : 8) * CTAS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc_g<CTAS>(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == PRODUCER_WARP) {
    // ------------------------------------------------------------------ TMA producer (one lane, every CTA)
    if (LS_ONE_LANE()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ready = 0;  // probe result for the current stage's `empty` barrier
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const uint32_t tx = leader ? (uint32_t)(stage_bytes * CTAS) : 0u;
#ifdef LS_GEMM_PROBE
      long long pr_wait = 0, pr_miss = 0;
      const long long pr_t0 = clock64();
#endif
      if (p.bres) {
        // Weight-resident mode (host: single CTA, no split-K, grid % n_tiles == 0 so that every tile of this CTA has the same
        // n-tile, single un-batched B): the B tile is loaded ONCE, the ring carries A k-blocks only.  A K = 320 linear
        // otherwise re-ingests the same 100 KB of weights for each of its M tiles - more than the 80 KB of A - and its
        // main loop runs at 840 clk per k-block instead of 583 (profiles/r2l_gemm_epilogue_experiments.txt (8)).
        const int nt0 = unit % p.n_tiles;
        const uint32_t bfull = smem_u32(b_full);
        mbar_expect_tx_u32(bfull, (uint32_t)bres_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_b_3d<1>(smem_base + kb * b_rows * 128, &p.mapB, bfull, kb * BK, nt0 * BN, 0);
        const uint32_t a_ring = smem_base + bres_bytes;
        for (int w = unit; w < total_work; w += nunits) {
          const int mu = w / p.n_tiles;
          int x0, y0, i0;
          decode_m_tile(p, mu, x0, y0, i0);
          int s = 0, tap = 0, cb = 0;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const int taps = p.seg_taps[s];
            const int dy = (taps == 9) ? (tap / 3 + p.tap9_off) : (taps == 4 ? (tap >> 1) + p.tap_dy0 : 0);
            const int dx = (taps == 9) ? (tap % 3 + p.tap9_off) : (taps == 4 ? (tap & 1) + p.tap_dx0 : 0);
            if (!ready) mbar_wait(&empty_bar[stage], phase ^ 1u);
            const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
            const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
            const uint32_t fb = full0 + stage * 8;
            ready = mbar_probe(empty0 + nstage * 8, nphase ^ 1u);
            mbar_expect_tx_u32(fb, (uint32_t)A_STAGE_BYTES);
            tma_a_4d<1>(a_ring + stage * A_STAGE_BYTES, &p.mapA[s], fb, cb * BK, p.cs * x0 + dx, p.cs * y0 + dy, i0);
            stage = nstage;
            phase = nphase;
            if (++cb == p.seg_cblk[s]) {
              cb = 0;
              if (++tap == taps) {
                tap = 0;
                ++s;
              }
            }
          }
        }
      } else
      for (int w = unit; w < total_work; w += nunits) {
        const int tile = w / p.splits;
        const int sp = w - tile * p.splits;
        const int kb0 = sp * p.kb_per;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per);
        const int mu = tile / p.n_tiles;
        const int nt = tile - mu * p.n_tiles;
        const int mt = mu * CTAS + rank;  // may run past m_tiles for the odd last pair: TMA zero-fills, stores clip
        int x0, y0, i0;
        decode_m_tile(p, mt, x0, y0, i0);
        const int bz = p.b_batched ? i0 : 0;
        const int brow = nt * BN + rank * b_rows;
        // position of k-block kb0 in the (segment, tap, channel block) nest
        int s = 0, rem = kb0;
        while (s + 1 < p.nseg && rem >= p.seg_taps[s] * p.seg_cblk[s]) {
          rem -= p.seg_taps[s] * p.seg_cblk[s];
          ++s;
        }
        int tap = rem / p.seg_cblk[s];
        int cb = rem - tap * p.seg_cblk[s];
        int kcol = kb0 * BK;
        if (p.kbs == 2) {
          // two k-blocks per stage: one probe, one arm, up to four loads
          for (int kb = kb0; kb < kb1; kb += 2) {
            const int n_here = min(2, kb1 - kb);
            if (!ready) mbar_wait(&empty_bar[stage], phase ^ 1u);
            const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
            const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t fb = full0 + stage * 8;
            ready = mbar_probe(empty0 + nstage * 8, nphase ^ 1u);
            if (leader) mbar_expect_tx_u32(fb, (uint32_t)(n_here * kblock_bytes * CTAS));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u < n_here) {
                const int taps = p.seg_taps[s];
                const int dy = (taps == 9) ? (tap / 3 + p.tap9_off) : (taps == 4 ? (tap >> 1) + p.tap_dy0 : 0);
                const int dx = (taps == 9) ? (tap % 3 + p.tap9_off) : (taps == 4 ? (tap & 1) + p.tap_dx0 : 0);
                tma_a_4d<CTAS>(sa + u * A_STAGE_BYTES, &p.mapA[s], fb, cb * BK, p.cs * x0 + dx, p.cs * y0 + dy, i0);
                tma_b_3d<CTAS>(sa + 2 * A_STAGE_BYTES + u * b_rows * 128, &p.mapB, fb, kcol, brow, bz);
                kcol += BK;
                if (++cb == p.seg_cblk[s]) {
                  cb = 0;
                  if (++tap == taps) {
                    tap = 0;
                    ++s;
                  }
                }
              }
            }
            stage = nstage;
            phase = nphase;
          }
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const int taps = p.seg_taps[s];
          const int dy = (taps == 9) ? (tap / 3 + p.tap9_off) : (taps == 4 ? (tap >> 1) + p.tap_dy0 : 0);
          const int dx = (taps == 9) ? (tap % 3 + p.tap9_off) : (taps == 4 ? (tap & 1) + p.tap_dx0 : 0);
#ifdef LS_GEMM_PROBE
          const long long pc0 = clock64();
          pr_miss += ready ? 0 : 1;
#endif
          if (!ready) mbar_wait(&empty_bar[stage], phase ^ 1u);
#ifdef LS_GEMM_PROBE
          pr_wait += clock64() - pc0;
#endif
          const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
          const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
          const uint32_t sa = smem_base + stage * stage_bytes;
#ifdef LS_GEMM_ABLATE
          if (CTAS == 1 && (p.flags & (LS_DBG_NO_A | LS_DBG_NO_B))) {
            // main-loop ablation (timing probes, tools/gemm_ablate.py; the result is garbage): drop one operand's load.
            // Same shape as produce_kblock (probe of the next stage first) so that only the TMA issue count differs.
            const uint32_t bytes = ((p.flags & LS_DBG_NO_A) ? 0u : (uint32_t)A_STAGE_BYTES) +
                                   ((p.flags & LS_DBG_NO_B) ? 0u : (uint32_t)(b_rows * 128));
            const uint32_t fb = full0 + stage * 8;
            uint32_t rdy;
            asm volatile(
                "{\n.reg .pred P;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
                "selp.u32 %0, 1, 0, P;\n}"
                : "=r"(rdy) : "r"(empty0 + nstage * 8), "r"(nphase ^ 1u) : "memory");
            if (bytes) mbar_expect_tx(&full_bar[stage], bytes); else mbar_arrive(&full_bar[stage]);
            if (!(p.flags & LS_DBG_NO_A))
              tma_load_4d(smem + stage * stage_bytes, &p.mapA[s], &full_bar[stage], cb * BK, p.cs * x0 + dx, p.cs * y0 + dy, i0);
            if (!(p.flags & LS_DBG_NO_B))
              tma_load_3d(smem + stage * stage_bytes + A_STAGE_BYTES, &p.mapB, &full_bar[stage], kcol, brow, bz);
            (void)fb;
            ready = rdy;
          } else
#endif
          ready = produce_kblock<CTAS>(sa, sa + A_STAGE_BYTES, &p.mapA[s], &p.mapB, full0 + stage * 8, tx, cb * BK,
                                       p.cs * x0 + dx, p.cs * y0 + dy, i0, kcol, brow, bz, empty0 + nstage * 8, nphase ^ 1u);
          kcol += BK;
          stage = nstage;
          phase = nphase;
          if (++cb == p.seg_cblk[s]) {
            cb = 0;
            if (++tap == taps) {
              tap = 0;
              ++s;
            }
          }
        }
      }
#ifdef LS_GEMM_PROBE
      if (blockIdx.x == 0)
        printf("gemm probe: producer total %lld clk, in wait(empty) %lld clk, probe misses %lld, kblocks/tile %d\n",
               clock64() - pr_t0, pr_wait, pr_miss, p.num_kb);
#endif
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer (one lane, pair leader only)
    if (leader && LS_ONE_LANE()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ready = 0;  // probe result for the current stage's `full` barrier
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      int lt = 0;
#ifdef LS_GEMM_PROBE
      long long mm_wait = 0, mm_miss = 0, mm_acc = 0;
      const long long mm_t0 = clock64();
#endif
      for (int w = unit; w < total_work; w += nunits, ++lt) {
        const int sp = w % p.splits;
        const int kb0 = sp * p.kb_per;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per);
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
#ifdef LS_GEMM_PROBE
        const long long ma0 = clock64();
#endif
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
#ifdef LS_GEMM_PROBE
        mm_acc += clock64() - ma0;
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        if (p.bres) {
          if (lt == 0) mbar_wait(b_full, 0u);
          const uint32_t a_ring = smem_base + bres_bytes;
          for (int kb = kb0; kb < kb1; ++kb) {
            if (!ready) mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
            const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
            ready = mbar_probe(full0 + nstage * 8, nphase);
            mma_quad<1>(d_tmem, umma_desc_sw128(a_ring + stage * A_STAGE_BYTES),
                        umma_desc_sw128(smem_base + kb * b_rows * 128), idesc, kb != kb0 ? 1u : 0u);
            umma_commit_g<1>(&empty_bar[stage]);
            stage = nstage;
            phase = nphase;
          }
          umma_commit_g<1>(&tmem_full[acc]);
          continue;
        }
        if (p.kbs == 2) {
          for (int kb = kb0; kb < kb1; kb += 2) {
            const int n_here = min(2, kb1 - kb);
            if (!ready) mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
            const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
            const uint32_t sa = smem_base + stage * stage_bytes;
            ready = mbar_probe(full0 + nstage * 8, nphase);
            mma_quad<CTAS>(d_tmem, umma_desc_sw128(sa), umma_desc_sw128(sa + 2 * A_STAGE_BYTES), idesc,
                           kb != kb0 ? 1u : 0u);
            if (n_here == 2)
              mma_quad<CTAS>(d_tmem, umma_desc_sw128(sa + A_STAGE_BYTES),
                             umma_desc_sw128(sa + 2 * A_STAGE_BYTES + b_rows * 128), idesc, 1u);
            umma_commit_g<CTAS>(&empty_bar[stage]);
            stage = nstage;
            phase = nphase;
          }
          umma_commit_g<CTAS>(&tmem_full[acc]);
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
#ifdef LS_GEMM_PROBE
          const long long mc0 = clock64();
          mm_miss += ready ? 0 : 1;
#endif
          if (!ready) mbar_wait(&full_bar[stage], phase);
#ifdef LS_GEMM_PROBE
          mm_wait += clock64() - mc0;
#endif
          tc_fence_after();
          const int nstage = (stage + 1 == stages) ? 0 : stage + 1;
          const uint32_t nphase = (nstage == 0) ? (phase ^ 1u) : phase;
          const uint32_t sa = smem_base + stage * stage_bytes;
#ifdef LS_GEMM_ABLATE
          if (CTAS == 1 && (p.flags & LS_DBG_NO_MMA)) {  // ablation: consume the stage without issuing MMAs
            uint32_t rdy;
            asm volatile(
                "{\n.reg .pred P;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
                "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%3];\n"
                "selp.u32 %0, 1, 0, P;\n}"
                : "=r"(rdy) : "r"(full0 + nstage * 8), "r"(nphase), "r"(empty0 + stage * 8) : "memory");
            ready = rdy;
          } else
#endif
          ready = mma_kblock<CTAS>(d_tmem, umma_desc_sw128(sa), umma_desc_sw128(sa + A_STAGE_BYTES), idesc,
                                   kb != kb0 ? 1u : 0u, empty0 + stage * 8, full0 + nstage * 8, nphase);
          stage = nstage;
          phase = nphase;
        }
        umma_commit_g<CTAS>(&tmem_full[acc]);  // accumulator complete once every MMA issued so far has retired
      }
#ifdef LS_GEMM_PROBE
      if (blockIdx.x == 0)
        printf("gemm probe: issuer total %lld clk (%d tiles), in wait(full) %lld clk, probe misses %lld, "
               "wait(tmem_empty) %lld clk\n", clock64() - mm_t0, lt, mm_wait, mm_miss, mm_acc);
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0..7, every CTA)
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int group = warp >> 2;  // 0: even slabs, 1: odd slabs
    const int r = q * 32 + lane;  // row of the tile == TMEM lane
    constexpr bool geglu = GEGLU;
    if (p.tma_store) {
      // Latency-tolerant epilogue: the tile's bias row goes to smem and ALL residual fragments of the group's slabs
      // are requested before waiting for the accumulator, so global-load latency hides behind the main loop.
      // Every warp stages and stores its own 32 rows of a slab (two 2 KB buffers, ping-pong): no block barrier and no
      // wait for the PREVIOUS store per slab.  (With one store per 128-row slab issued behind a 128-thread barrier
      // and a wait on the preceding store, each slab cost ~1500 clocks of latency: a K = 64 GEMM took as long as K = 320.)
      uint8_t* my_stage = staging + warp * (p.slab_single ? WSLAB_BYTES : 2 * WSLAB_BYTES);  // 32 rows x 128 (64) bytes
      constexpr int NSLAB_MAX = 4;  // split-K parking: slabs per group at BN = 256
      const int nslab = geglu ? BN / 64 : BN / 32;
      const int n_out_total = geglu ? p.N / 2 : p.N;
      const int gtid = (warp - 4 * group) * 32 + lane;  // 0..127 inside the group
      int lt = 0;
      // bias row of the NEXT tile is fetched into registers one tile ahead and parked in a double-buffered smem row, so
      // its global-load latency never sits on the epilogue's critical path (ncu: 9 % of the stall samples before)
      float bias_next[2] = {0.f, 0.f};
      float cs_next[2] = {0.f, 0.f};  // LN == 1: col_sum of the next tile's columns, parked like the bias row
      auto fetch_bias = [&](int w_next) {
        bias_next[0] = bias_next[1] = 0.f;
        if constexpr (LN == 1) cs_next[0] = cs_next[1] = 0.f;
        if (w_next >= total_work || (LN != 1 && p.bias == nullptr)) return;
        const int tile = (int)fdiv((uint32_t)w_next, p.fd_splits);
        const int mu = (int)fdiv((uint32_t)tile, p.fd_n_tiles);
        const int nt = tile - mu * p.n_tiles;
        int x0, y0, i0;
        decode_m_tile(p, mu * CTAS + rank, x0, y0, i0);
        const int64_t m0 = ((int64_t)i0 * p.H + y0) * p.W + x0;
        if constexpr (LN == 1) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int c = gtid + 128 * k;
            cs_next[k] = (c < BN && nt * BN + c < p.N) ? __ldg(p.colsum + nt * BN + c) : 0.f;
          }
          // the next tile's row partials -> L1 now, so that their L2 latency (two dependent rounds of loads: ~1400 clocks
          // per tile on launches that are epilogue-bound anyway) is not paid at the top of that tile
          if (m0 + r < p.M)
            for (int i = 0; i < p.ln_nparts_in; ++i)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p.ln_in + (int64_t)i * p.ln_in_stride + m0 + r));
          if (p.bias == nullptr) return;
        }
        const int64_t brow_i = (p.bias_div > 0 && m0 < p.M) ? (int64_t)fdiv((uint32_t)m0, p.fd_bias_div) : 0;  // M < 2^31 here
        const float* brow = p.bias + brow_i * (int64_t)p.bias_ld + nt * BN;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = gtid + 128 * k;
          bias_next[k] = (c < BN && nt * BN + c < p.N) ? __ldg(brow + c) : 0.f;
        }
      };
      // With >= 3 tiles per CTA (host: epi_alt) the two groups take ALTERNATE tiles (group g owns accumulator g): the
      // per-tile serial part (decode, bias row, residual prefetch, barriers: ~1700 clocks) of one group overlaps the slab
      // work of the other (level-0 short-K linear 18.0 -> 16.5 us).  With one or two tiles per CTA both groups share
      // every tile (contiguous halves of its slabs), which finishes a lone tile sooner.
      const bool alt = (p.epi_alt != 0);
      const int w_first = unit + (alt ? group * nunits : 0);
      const int w_step = alt ? 2 * nunits : nunits;
      fetch_bias(w_first);
#ifdef LS_GEMM_PROBE
      long long ep[5][12];
      for (int i = 0; i < 5; ++i)
        for (int k = 0; k < 12; ++k) ep[i][k] = 0;
      const long long ep_t0 = clock64();
#define EP_STAMP(k) do { if (lt < 4) ep[lt][k] = clock64() - ep_t0; } while (0)
#else
#define EP_STAMP(k) do {} while (0)
#endif
      lt = alt ? group : 0;
      for (int w = w_first; w < total_work; w += w_step, lt += (alt ? 2 : 1)) {
        const int tile = (int)fdiv((uint32_t)w, p.fd_splits);
        const int sp = w - tile * p.splits;
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
        const int mu = (int)fdiv((uint32_t)tile, p.fd_n_tiles);
        const int nt = tile - mu * p.n_tiles;
        const int mt = mu * CTAS + rank;
        int x0, y0, i0;
        decode_m_tile(p, mt, x0, y0, i0);
        const int64_t m0 = ((int64_t)i0 * p.H + y0) * p.W + x0;  // tile rows are contiguous (checked on the host)
        const int64_t m = m0 + r;
        const bool row_ok = m < p.M;
        const bool tile_ok = m0 < p.M;
        EP_STAMP(8);
        const int g0 = (group + lt) & 1;  // split-K parking: slab parity of this group alternates per tile
        // normal path: the group's slabs [j_lo, j_hi) are stored in pairs from j_lo on, an odd count leaves a single
        // shared tile: with an odd slab count the larger half alternates between the groups (5 = 3 + 2, then 2 + 3)
        const int n_first = ((lt & 1) == 0) ? (nslab + 1) / 2 : nslab / 2;
        const int j_lo = alt ? 0 : (group == 0 ? 0 : n_first);
        const int j_hi = alt ? nslab : (group == 0 ? n_first : nslab);
        const int n_mine = j_hi - j_lo;
        // (1) this tile's bias row (host guarantees one row per tile: bias_div % 128 == 0) -> the group's smem row.  Its
        //     previous contents were last read before the final slab barrier of the previous tile.
        float* my_bias = bias_sm + group * 256;
        my_bias[gtid] = bias_next[0];
        my_bias[gtid + 128] = bias_next[1];
        // LN == 1: out = rs_a * acc + rs_c * col_sum[n] + bias[n]
        float* my_cs = colsum_sm + group * 256;
        float rs_a = 1.f, rs_c = 0.f;
        if constexpr (LN == 1) {
          my_cs[gtid] = cs_next[0];
          my_cs[gtid + 128] = cs_next[1];
          if (row_ok) {
            // the producer's partial sums of this row, combined in part order (deterministic).  Six loads are issued
            // together, then added in order (as a plain loop ptxas emits load -> add -> load: one latency per part,
            // ~360 clocks each at the top of every tile; absent parts add an exact 0)
            const float2* pp = p.ln_in + m;
            float su = 0.f, sq = 0.f;
            for (int i0 = 0; i0 < p.ln_nparts_in; i0 += 6) {
              float2 t[6];
#pragma unroll
              for (int k = 0; k < 6; ++k)
                t[k] = (i0 + k < p.ln_nparts_in) ? __ldg(pp + (int64_t)(i0 + k) * p.ln_in_stride) : make_float2(0.f, 0.f);
#pragma unroll
              for (int k = 0; k < 6; ++k) {
                su += t[k].x;
                sq += t[k].y;
              }
            }
            const float mean = su * p.ln_inv_k;
            const float rstd = rsqrtf(fmaxf(sq * p.ln_inv_k - mean * mean, 0.f) + p.ln_eps);
            rs_a = rstd;
            rs_c = -rstd * mean;
          }
        }
        // LN == 2: this thread's row over three FIXED slab ranges of the tile, [0, nslab/2), [nslab/2, (nslab+1)/2),
        // [(nslab+1)/2, nslab) - the only boundaries the groups ever split a tile at - so that the partials (and with them
        // every bit downstream) do not depend on which group, or how many tiles per CTA, a launch geometry happens to use
        float ln_su[3] = {0.f, 0.f, 0.f}, ln_sq[3] = {0.f, 0.f, 0.f};
        const int ln_lo = nslab / 2, ln_hi = (nslab + 1) / 2;
        // (2) residual of the NEXT pair of slabs, requested one pair ahead with COALESCED loads: instruction i of a
        //     pair (32 rows x 128 bytes per warp) has lane l read 16-byte chunk (l & 7) of row 4 i + (l >> 3), i.e. four
        //     whole 128-byte lines per instruction.  (A thread reading its own row's 64 bytes - the layout the
        //     accumulator arrives in - touches 32 lines per instruction: the L1 spent ~2000 wavefronts per tile on it,
        //     as much as the whole main loop of a K = 320 linear.)  The chunks reach their rows through the warp's
        //     staging buffer: stored at the 128B-swizzled position the output chunk will take, read back by the row's
        //     thread (both patterns conflict-free), then overwritten by the output.
        uint4 res[8];
        const bool has_res = (LN != 1) && (p.residual != nullptr);  // a folded-LayerNorm consumer never adds a residual (host)
        auto fetch_res_pair = [&](int s0) {  // slabs j_lo + s0, j_lo + s0 + 1 (or the trailing single slab)
          const int j = j_lo + s0;
          const int ncol0 = (geglu ? nt * (BN / 2) : nt * BN) + j * 32;  // output column, as in the slab loop below
          const bool single = (j == j_hi - 1);
          const int64_t mrow0 = m0 + q * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // pair: 4 rows x 8 chunks per instruction; single slab (64-byte rows): 8 rows x 4 chunks, 4 instructions
            const int rr = single ? (i * 8 + (lane >> 2)) : (i * 4 + (lane >> 3));
            const int cc = single ? (lane & 3) : (lane & 7);
            const int col = ncol0 + cc * 8;
            const bool ok = has_res && j < j_hi && (!single || i < 4) && (mrow0 + rr < p.M) && (col + 8 <= n_out_total);
            res[i] = ok ? __ldg(reinterpret_cast<const uint4*>(p.residual + (mrow0 + rr) * (int64_t)p.ldr + col))
                        : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        EP_STAMP(9);
        if (has_res) fetch_res_pair(0);
        EP_STAMP(10);
        fetch_bias(w + w_step);  // (3) this group's next work item's bias row -> registers
        EP_STAMP(0);
        named_bar_sync(1 + group, 128);  // bias row visible to the group
        EP_STAMP(1);
        mbar_wait_relaxed(&tmem_full[acc], acc_phase);
        tc_fence_after();
        EP_STAMP(2);
        const uint32_t tbase = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS;
        bool released = false;
        if (p.splits > 1) {
          // split-K.  Phase A: park this CTA's fp32 partial in the workspace and hand the accumulator back.
#pragma unroll
          for (int s = 0; s < NSLAB_MAX; ++s) {
            const int j = g0 + 2 * s;
            const int ncol0 = nt * BN + j * 32;
            if (j >= nslab || ncol0 >= n_out_total) break;
            uint32_t v[32];
            tmem_ld_32x32(tbase + j * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              float* dst = p.ws + ((int64_t)sp * p.M + m) * p.N + ncol0;
              if (ncol0 + 32 <= n_out_total) {
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4)
                  __stcg(reinterpret_cast<float4*>(dst) + e4,
                         make_float4(__uint_as_float(v[e4 * 4]), __uint_as_float(v[e4 * 4 + 1]),
                                     __uint_as_float(v[e4 * 4 + 2]), __uint_as_float(v[e4 * 4 + 3])));
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                  if (ncol0 + e < n_out_total) __stcg(dst + e, __uint_as_float(v[e]));
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
          __threadfence();
          named_bar_sync(3, 256);
          // Rendezvous of the `splits` CTAs that share this tile.  They are all resident: the host only splits when the
          // whole launch is a single wave of one CTA per SM.
          unsigned int* tk = p.tickets + 2 * tile;
          if (warp == 0 && lane == 0) {
            atomicAdd(tk, 1u);
            uint32_t spins = 0;
            while (*reinterpret_cast<volatile unsigned int*>(tk) < (unsigned int)p.splits) {
              __nanosleep(64);
              if (++spins > (1u << 24)) {
                printf("latentsync_b200: split-K rendezvous timeout (block %d)\n", blockIdx.x);
                __trap();
              }
            }
          }
          named_bar_sync(3, 256);
          __threadfence();
          // Phase B: every CTA finishes 1/splits of the tile's 8-column chunks: sum of the partials in split order
          // (deterministic) + bias + residual -> fp16, coalesced 16-byte accesses.
          {
            const int ncols = min(BN, p.N - nt * BN);
            const int cpr = ncols >> 3;
            const int total = BM * cpr;
            const int per = (total + p.splits - 1) / p.splits;
            const int i1 = min(total, (sp + 1) * per);
            const int64_t brow_i = (p.bias_div > 0 && tile_ok) ? (m0 / p.bias_div) : 0;
            const float* brow = p.bias != nullptr ? p.bias + brow_i * (int64_t)p.bias_ld : nullptr;
            __half* outp = reinterpret_cast<__half*>(p.out);
            for (int i = sp * per + warp * 32 + lane; i < i1; i += 256) {
              const int row = i / cpr;
              const int col = nt * BN + (i - row * cpr) * 8;
              const int64_t mm = m0 + row;
              if (mm >= p.M) continue;
              float a8[8];
              if (brow != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(brow + col + 4));
                a8[0] = b0.x; a8[1] = b0.y; a8[2] = b0.z; a8[3] = b0.w;
                a8[4] = b1.x; a8[5] = b1.y; a8[6] = b1.z; a8[7] = b1.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) a8[e] = 0.f;
              }
              uint4 rs = make_uint4(0u, 0u, 0u, 0u);
              if (p.residual != nullptr) rs = __ldg(reinterpret_cast<const uint4*>(p.residual + mm * (int64_t)p.ldr + col));
              const float* src = p.ws + mm * (int64_t)p.N + col;
              const int64_t sstride = p.M * (int64_t)p.N;
#pragma unroll 4
              for (int q2 = 0; q2 < p.splits; ++q2) {
                const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + q2 * sstride));
                const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + q2 * sstride + 4));
                a8[0] += t0.x; a8[1] += t0.y; a8[2] += t0.z; a8[3] += t0.w;
                a8[4] += t1.x; a8[5] += t1.y; a8[6] += t1.z; a8[7] += t1.w;
              }
              const __half2* rh = reinterpret_cast<const __half2*>(&rs);
              uint4 o;
              __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = __half22float2(rh[e]);
                oh[e] = __floats2half2_rn(a8[2 * e] + t.x, a8[2 * e + 1] + t.y);
              }
              *reinterpret_cast<uint4*>(outp + mm * (int64_t)p.ldo + col) = o;
            }
          }
          named_bar_sync(3, 256);
          if (warp == 0 && lane == 0) {
            const unsigned int old = atomicAdd(tk + 1, 1u);
            if (old == (unsigned int)p.splits - 1u) {  // last CTA to leave resets both counters for the next launch
              tk[0] = 0u;
              tk[1] = 0u;
            }
          }
          continue;
        }
        // Normal path.  The group owns a CONTIGUOUS range of slab pairs (the larger half alternates between the groups
        // from tile to tile); a pair (64 output columns = one 128-byte line per row) is staged in the warp's own 4 KB
        // buffer (128-byte swizzle) and leaves with ONE TMA store per warp.  Measured (clock64 probes, K = 64 so that
        // only the epilogue counts): a store costs its issuer ~450 clocks whatever the box, so 32-column boxes
        // (64-byte rows) made the TMA unit the bound of every short-K launch at ~16 B/clk/SM; coalesced st.global
        // from the staging buffer was slower still (15.0 vs 11.5 us at M = 32768, N = 320).
#pragma unroll 1
        for (int pi = 0; 2 * pi < n_mine; ++pi) {
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const int s = 2 * pi + hs;
          const int j = j_lo + s;
          if (j >= j_hi) break;
          const int ncol0 = geglu ? nt * (BN / 2) + j * 32 : nt * BN + j * 32;  // first output column of the slab
          if (ncol0 >= n_out_total) break;  // this and all later slabs lie beyond N (uniform over the group)
          uint32_t v[32];
          float f[32];
          if (!geglu) {
            tmem_ld_32x32(tbase + j * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 t = *reinterpret_cast<const float4*>(my_bias + j * 32 + e4 * 4);
              if constexpr (LN == 1) {
                const float4 c4 = *reinterpret_cast<const float4*>(my_cs + j * 32 + e4 * 4);
                f[e4 * 4] = fmaf(rs_a, __uint_as_float(v[e4 * 4]), fmaf(rs_c, c4.x, t.x));
                f[e4 * 4 + 1] = fmaf(rs_a, __uint_as_float(v[e4 * 4 + 1]), fmaf(rs_c, c4.y, t.y));
                f[e4 * 4 + 2] = fmaf(rs_a, __uint_as_float(v[e4 * 4 + 2]), fmaf(rs_c, c4.z, t.z));
                f[e4 * 4 + 3] = fmaf(rs_a, __uint_as_float(v[e4 * 4 + 3]), fmaf(rs_c, c4.w, t.w));
              } else {
                f[e4 * 4] = __uint_as_float(v[e4 * 4]) + t.x;
                f[e4 * 4 + 1] = __uint_as_float(v[e4 * 4 + 1]) + t.y;
                f[e4 * 4 + 2] = __uint_as_float(v[e4 * 4 + 2]) + t.z;
                f[e4 * 4 + 3] = __uint_as_float(v[e4 * 4 + 3]) + t.w;
              }
            }
          } else {
            uint32_t g[32];
            tmem_ld_32x32(tbase + j * 32, v);
            tmem_ld_32x32(tbase + BN / 2 + j * 32, g);
            tmem_ld_wait();
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              float4 bv = *reinterpret_cast<const float4*>(my_bias + j * 32 + e4 * 4);
              float4 bg = *reinterpret_cast<const float4*>(my_bias + BN / 2 + j * 32 + e4 * 4);
              if constexpr (LN == 1) {  // LayerNorm in front of the GEGLU projection: fix up value and gate pre-activations
                const float4 cv = *reinterpret_cast<const float4*>(my_cs + j * 32 + e4 * 4);
                const float4 cg = *reinterpret_cast<const float4*>(my_cs + BN / 2 + j * 32 + e4 * 4);
                bv.x = fmaf(rs_c, cv.x, bv.x); bv.y = fmaf(rs_c, cv.y, bv.y);
                bv.z = fmaf(rs_c, cv.z, bv.z); bv.w = fmaf(rs_c, cv.w, bv.w);
                bg.x = fmaf(rs_c, cg.x, bg.x); bg.y = fmaf(rs_c, cg.y, bg.y);
                bg.z = fmaf(rs_c, cg.z, bg.z); bg.w = fmaf(rs_c, cg.w, bg.w);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[e4 * 4 + e] = __float_as_uint(rs_a * __uint_as_float(v[e4 * 4 + e]));
                  g[e4 * 4 + e] = __float_as_uint(rs_a * __uint_as_float(g[e4 * 4 + e]));
                }
              }
              f[e4 * 4] = (__uint_as_float(v[e4 * 4]) + bv.x) * gelu_erf_f(__uint_as_float(g[e4 * 4]) + bg.x);
              f[e4 * 4 + 1] =
                  (__uint_as_float(v[e4 * 4 + 1]) + bv.y) * gelu_erf_f(__uint_as_float(g[e4 * 4 + 1]) + bg.y);
              f[e4 * 4 + 2] =
                  (__uint_as_float(v[e4 * 4 + 2]) + bv.z) * gelu_erf_f(__uint_as_float(g[e4 * 4 + 2]) + bg.z);
              f[e4 * 4 + 3] =
                  (__uint_as_float(v[e4 * 4 + 3]) + bv.w) * gelu_erf_f(__uint_as_float(g[e4 * 4 + 3]) + bg.w);
            }
          }
          if (j == j_hi - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
            released = true;
          }
          // stage address of this thread's row: pair = 128-byte rows, chunk c at c ^ (lane & 7) (SWIZZLE_128B);
          // trailing single slab = 64-byte rows, chunk c at c ^ ((lane >> 1) & 3) (SWIZZLE_64B)
          const int half_sel = hs;
          // slab_single launches (main-loop-bound, no residual: the 16 KB of staging saved buy one more pipeline stage)
          // stage and store EVERY slab on its own, like the trailing slab of an odd count
          const bool single = p.slab_single || ((half_sel == 0) && (s == n_mine - 1));
          if (has_res && hs == 0) {
            // the previous store has finished reading the buffer -> park this pair's residual chunks, request the next pair's
            const bool single_p = single;
            if (lane == 0) bulk_wait_group_read<0>();
            __syncwarp();
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (single_p) {
                if (i < 4) {
                  const int rr = i * 8 + (lane >> 2), cc = lane & 3;
                  *reinterpret_cast<uint4*>(my_stage + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4)) = res[i];
                }
              } else {
                const int rr = i * 4 + (lane >> 3), cc = lane & 7;
                *reinterpret_cast<uint4*>(my_stage + rr * 128 + ((cc ^ (rr & 7)) << 4)) = res[i];
              }
            }
            __syncwarp();
            fetch_res_pair(2 * pi + 2);
          }
          if (has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint8_t* src = single ? my_stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)
                                          : my_stage + lane * 128 + (((half_sel * 4 + c) ^ (lane & 7)) << 4);
              const uint4 u = *reinterpret_cast<const uint4*>(src);
              const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = __half22float2(h2[e]);
                f[c * 8 + e * 2] += t.x;
                f[c * 8 + e * 2 + 1] += t.y;
              }
            }
          }
          if (p.flags & LS_EPI_SILU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = silu_f(f[e]);
          }
          // stage.  Pair: row `lane` of the warp's buffer is 128 bytes, 16-byte chunk c sits at c ^ (lane & 7)
          // (SWIZZLE_128B).  Trailing single slab: 64-byte rows, chunk c at c ^ ((lane >> 1) & 3) (SWIZZLE_64B).
          if ((half_sel == 0 || p.slab_single) && !has_res) {  // the previous store has finished reading the buffer
            if (lane == 0) bulk_wait_group_read<0>();
            __syncwarp();
          }
          float slab_su = 0.f, slab_sq = 0.f, slab_su2 = 0.f, slab_sq2 = 0.f;  // two chains each: ILP
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[c * 8 + e * 2], f[c * 8 + e * 2 + 1]);
            if constexpr (LN == 2) {
              // statistics of the fp32 values before their rounding to fp16 (two instructions per element; from the
              // rounded halves it is three and a half, and these launches are bound by the epilogue's issue slots)
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                slab_su += f[c * 8 + e];
                slab_su2 += f[c * 8 + e + 1];
                slab_sq = fmaf(f[c * 8 + e], f[c * 8 + e], slab_sq);
                slab_sq2 = fmaf(f[c * 8 + e + 1], f[c * 8 + e + 1], slab_sq2);
              }
            }
            uint8_t* dst = single ? my_stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)
                                  : my_stage + lane * 128 + (((half_sel * 4 + c) ^ (lane & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = u;
          }
          if constexpr (LN == 2) {
            const int pid = (j >= ln_lo ? 1 : 0) + (j >= ln_hi ? 1 : 0);
            slab_su += slab_su2;
            slab_sq += slab_sq2;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              ln_su[k] += (pid == k) ? slab_su : 0.f;
              ln_sq[k] += (pid == k) ? slab_sq : 0.f;
            }
          }
          EP_STAMP(3 + (s & 3));
          if (half_sel == 1 || single || ncol0 + 32 >= n_out_total) {
            // staged: one TMA store of the warp's 32 rows x 64 (32) columns; columns >= N and rows >= M are clipped
            // by the TMA unit.  128-byte rows: half as many row requests per byte as a 32-column box.
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && m0 + q * 32 < p.M) {
              const int col0 = single ? ncol0 : ncol0 - half_sel * 32;
              if (p.up2) {
                // the warp's 32 rows = up2_bx x up2_by low-resolution pixels of the tile's box -> every second pixel / row
                const int o = q * 32;  // first row of the warp inside the tile: x fastest, then y, then image
                const int ox = o % p.bw, oy = (o / p.bw) % p.bh, on = o / (p.bw * p.bh);
                tma_store_4d(single ? &p.mapOut32 : &p.mapOut, my_stage, col0, x0 + ox, y0 + oy, i0 + on);
              } else {
                tma_store_2d(single ? &p.mapOut32 : &p.mapOut, my_stage, col0, (int)(m0 + q * 32));
              }
              bulk_commit_group();
            }
            if constexpr (LN == 3) {
              // column sums of what was just staged (the store only reads the buffer) -> this quarter's row of gn_sm.
              // Host: N % BN == 0, so a pair is never cut short by the end of N.
              float2* grow = gn_sm + (size_t)((alt ? group : 0) * 4 + q) * BN;
              if (single) gn_colsum_single(my_stage, lane, grow + j * 32);
              else gn_colsum_pair(my_stage, lane, grow + (j - 1) * 32);
            }
          }
        }
        }
        EP_STAMP(7);
        if constexpr (LN == 2) {
          if (row_ok) {  // 32 consecutive rows per warp: one coalesced 256-byte store per part
            // every part is written exactly once per tile, by the group whose slab range contains it (zeros where the
            // range is empty or lies beyond N)
            const bool own[3] = {alt || group == 0,
                                 alt || (group == 0 ? (n_first == ln_hi) : (n_first == ln_lo && ln_lo != ln_hi)),
                                 alt || group == 1};
            float2* dst = p.ln_out + (int64_t)(3 * nt) * p.ln_out_stride + m;
#pragma unroll
            for (int k = 0; k < 3; ++k)
              if (own[k]) dst[(int64_t)k * p.ln_out_stride] = make_float2(ln_su[k], ln_sq[k]);
          }
        }
        if constexpr (LN == 3) {
          // tile-level reduction of the column sums: quarters in order, then the columns of each unit in order.  With
          // alternate tiles a group owns its tile (and its slot of gn_sm); otherwise both groups share the tile.
          const int nthr = alt ? 128 : 256;
          const int tid = alt ? gtid : warp * 32 + lane;
          float2* base = gn_sm + (size_t)(alt ? group : 0) * 4 * BN;
          if (alt) named_bar_sync(1 + group, 128); else named_bar_sync(3, 256);
          for (int c = tid; c < BN; c += nthr) {
            const float2 a = base[c], b = base[BN + c], c2 = base[2 * BN + c], d = base[3 * BN + c];
            base[c] = make_float2(((a.x + b.x) + c2.x) + d.x, ((a.y + b.y) + c2.y) + d.y);
          }
          if (alt) named_bar_sync(1 + group, 128); else named_bar_sync(3, 256);
          if (tid < p.gn_upt && tile_ok) {
            float su = 0.f, sq = 0.f;
            const float2* src = base + tid * p.gn_unit;
            for (int i = 0; i < p.gn_unit; ++i) {
              su += src[i].x;
              sq += src[i].y;
            }
            p.gn_out[(m0 >> 7) * (int64_t)p.gn_ld + nt * p.gn_upt + tid] = make_float2(su, sq);
          }
          if (!alt) named_bar_sync(3, 256);  // the other group must not start the next tile's sums before these reads
        }
        named_bar_sync(1 + group, 128);  // every warp of the group has read the bias row: the next tile may overwrite it
        if (!released) {  // group had no slab inside N for this tile: still hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
        }
      }
#ifdef LS_GEMM_PROBE
      if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 3))
        for (int i = 0; i < 4; ++i)
          printf("gemm probe: epilogue warp %d tile %d: decoded %lld bias_st %lld res %lld top %lld bias_bar %lld acc_ready %lld slabs %lld %lld %lld %lld stored %lld\n",
                 warp, i, ep[i][8], ep[i][9], ep[i][10], ep[i][0], ep[i][1], ep[i][2], ep[i][3], ep[i][4], ep[i][5], ep[i][6], ep[i][7]);
#endif
    } else if (group == 0) {
      // legacy path (fp32 output / ragged geometry / narrow leading dimension): per-thread direct global stores
      const int ix = r % p.bw;
      const int iy = (r / p.bw) % p.bh;
      const int in = r / (p.bw * p.bh);
      int lt = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, ++lt) {  // splits == 1 on this path
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
        const int mu = tile / p.n_tiles;
        const int nt = tile - mu * p.n_tiles;
        const int mt = mu * CTAS + rank;
        int x0, y0, i0;
        decode_m_tile(p, mt, x0, y0, i0);
        const int x = x0 + ix, y = y0 + iy, img = i0 + in;
        const bool row_ok = (x < p.W) && (y < p.H) && (img < p.nimg);
        const int64_t m = ((int64_t)img * p.H + y) * p.W + x;

        mbar_wait_relaxed(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS;
        if (!geglu) {
          const int nchunk = BN / 32;
#pragma unroll 1
          for (int c = 0; c < nchunk; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tbase + c * 32, v);
            tmem_ld_wait();
            if (c == nchunk - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
            }
            const int n_base = nt * BN + c * 32;
            if (row_ok && n_base < p.N) {
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              add_bias32(p, m, n_base, f);
              epilogue_store32(p, m, n_base, p.N, f);
            }
          }
        } else {
          // W rows of this tile are [BN/2 value rows | BN/2 gate rows]
          const int nchunk = BN / 64;
#pragma unroll 1
          for (int c = 0; c < nchunk; ++c) {
            uint32_t v[32], g[32];
            tmem_ld_32x32(tbase + c * 32, v);
            tmem_ld_32x32(tbase + BN / 2 + c * 32, g);
            tmem_ld_wait();
            if (c == nchunk - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
            }
            const int nv_base = nt * BN + c * 32;  // packed column of the value half
            if (row_ok && nv_base < p.N) {
              float fv[32], fg[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                fv[j] = __uint_as_float(v[j]);
                fg[j] = __uint_as_float(g[j]);
              }
              add_bias32(p, m, nv_base, fv);
              add_bias32(p, m, nv_base + BN / 2, fg);
#pragma unroll
              for (int j = 0; j < 32; ++j) fv[j] = fv[j] * gelu_erf_f(fg[j]);
              epilogue_store32(p, m, nt * (BN / 2) + c * 32, p.N / 2, fv);
            }
          }
        }
      }
    } else {
      // the legacy path uses the first 4 epilogue warps only; the others just hand the accumulators back
      int lt = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
        mbar_wait_relaxed(&tmem_full[acc], acc_phase);
        if (lane == 0) mbar_arrive_leader<CTAS>(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == MMA_WARP) tmem_dealloc_g<CTAS>(tmem_base, TMEM_COLS);
}

template <bool GEGLU, int LN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  gemm_body<1, GEGLU, LN>(p);
}
template <bool GEGLU, int LN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_pair_kernel(const __grid_constant__ GemmKParams p) {
  gemm_body<2, GEGLU, LN>(p);
}
typedef void (*GemmKernelFn)(const GemmKParams);
// [CTAS - 1][geglu][LN]; a GEGLU projection never produces LayerNorm / GroupNorm partials (its output feeds a Linear)
static GemmKernelFn gemm_kernel_of(int ctas, bool geglu, int ln) {
  static const GemmKernelFn tab[2][2][4] = {
      {{gemm_tc_kernel<false, 0>, gemm_tc_kernel<false, 1>, gemm_tc_kernel<false, 2>, gemm_tc_kernel<false, 3>},
       {gemm_tc_kernel<true, 0>, gemm_tc_kernel<true, 1>, nullptr, nullptr}},
      {{gemm_tc_pair_kernel<false, 0>, gemm_tc_pair_kernel<false, 1>, gemm_tc_pair_kernel<false, 2>,
        gemm_tc_pair_kernel<false, 3>},
       {gemm_tc_pair_kernel<true, 0>, gemm_tc_pair_kernel<true, 1>, nullptr, nullptr}}};
  return tab[ctas - 1][geglu ? 1 : 0][ln];
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

static int num_sms() {
  static int n_dev[16] = {};  // per device: a process may drive several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return 148;
  int& n = n_dev[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Estimated cycles of one launch: waves x max(main loop, epilogue) + tail, from MEASURED clocks per 64-wide k-block
// (profiles/r2_gemm_ablate_elect.txt, tools/gemm_ablate.py on B200, both control loops under elect.sync):
//   single CTA, 128 x BN tile : 434 / 465 / 583 / 998 clk at BN = 64 / 128 / 160 / 256 (tensor work: 2 BN clk)
//   CTA pair,   256 x BN tile : 392 / 397 / 502 / 637 / 885 clk at BN = 64 / 128 / 160 / 192 / 256 PER PAIR - i.e. half that
//                               per 128 rows: each CTA stages only half of the B tile (cta_group::2)
// and checked against the sweep of the UNet's own shapes (profiles/r2_gemm_shapes_elect.txt, tools/gemm_shapes.py): the
// model's choice is within 2 % of the best measured configuration summed over those shapes.  Pairs win wherever the
// main loop dominates (3x3 convolutions, K >= 2048 linears: 256 x 160 tiles, +9-14 %); they lose on short-K launches
// (epilogue-bound, and a pair launch costs ~2 us more: cluster scheduling + two cluster barriers).
static double tile_cost(int ctas, int bn, int splits, int m_tiles, int N, int num_kb, int sms, bool geglu, bool res) {
  // (the ablation's N = 18 tiles / M = 2048 problem is the L2-contention worst case for wide tiles; the wide-tile entries
  // below are the per-k-block clocks of the large-M VAE convolutions, profiles/r2_gemm_shapes_vae.txt)
  static const double T1[9] = {0, 420, 434, 450, 495, 583, 680, 780, 880};
  static const double T2[9] = {0, 392, 400, 430, 460, 505, 600, 700, 815};
  const int n_tiles = (N + bn - 1) / bn;
  const int m_units = (m_tiles + ctas - 1) / ctas;
  const long work = (long)m_units * n_tiles * splits;
  const int units = sms / ctas;
  const long waves = (work + units - 1) / units;
  const double t_kb = (ctas == 1 ? T1 : T2)[bn / 32];
  const int nslab = geglu ? bn / 64 : bn / 32;
  double t_epi = 1200.0 + nslab * (geglu ? 1500.0 : (res ? 1000.0 : 800.0));
  if (ctas == 2) t_epi *= 1.05;
  const int kb_per = (num_kb + splits - 1) / splits;
  const double t_main = kb_per * t_kb;
  const double t_tile = (t_main > t_epi ? t_main : t_epi);
  // split-K: every CTA parks an fp32 partial, meets its peers and finishes 1/splits of the tile
  // measured (tools/l3_probe.py, cold weights): the level-3 conv (K = 11520) takes 35 / 30 / 57 us at 2 / 4 / 8 splits -
  // beyond 4 the fp32 partials (splits x M x N x 4 B written and read back) cost more than the shorter K loop saves
  const double t_split = splits > 1 ? 6000.0 + 300.0 * splits + (splits > 4 ? 12000.0 : 0.0) : 0.0;
  return waves * t_tile + t_epi + t_split + 3000.0 + (ctas == 2 ? 4000.0 : 0.0);
}

// split-K scratch (fp32 partials + per-tile tickets), owned by the library, grown outside graph capture
struct SplitScratch {
  ScratchBlock ws_b, tickets_b;
};
static SplitScratch g_split[16];

static int gemm_impl(const LsGemmArgs* a, cudaStream_t stream) {
  LS_CHECK(a != nullptr, "ls_gemm: null args");
  LS_CHECK(a->nseg >= 1 && a->nseg <= LS_GEMM_MAX_SEG, "ls_gemm: nseg=%d out of range", a->nseg);
  LS_CHECK(a->nimg >= 1 && a->H >= 1 && a->W >= 1, "ls_gemm: bad geometry");
  LS_CHECK(a->N >= 1 && a->out != nullptr && a->b_ptr != nullptr, "ls_gemm: bad output/weights");
  PFN_tmapEncodeTiled encode = get_encode_fn();
  LS_CHECK(encode != nullptr, "ls_gemm: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = a->nseg;
  p.H = a->H;
  p.W = a->W;
  p.nimg = a->nimg;
  bool any_conv = false;
  int ktot = 0;
  for (int s = 0; s < a->nseg; ++s) {
    LS_CHECK(a->a_ptr[s] != nullptr, "ls_gemm: segment %d null", s);
    LS_CHECK(a->a_ch[s] > 0 && a->a_ch[s] % BK == 0, "ls_gemm: segment %d channels %d not a multiple of 64", s,
             a->a_ch[s]);
    LS_CHECK(a->a_ld[s] >= a->a_ch[s] && a->a_ld[s] % 8 == 0, "ls_gemm: segment %d ld %d invalid", s, a->a_ld[s]);
    LS_CHECK(a->a_taps[s] == 1 || a->a_taps[s] == 9 || (a->a_taps[s] == 4 && a->up2 >= 1 && a->up2 <= 4),
             "ls_gemm: taps must be 1 or 9 (or 4 for a sub-pixel phase, up2 = 1..4)");
    LS_CHECK((reinterpret_cast<uintptr_t>(a->a_ptr[s]) & 15) == 0, "ls_gemm: segment %d pointer not 16B aligned", s);
    any_conv |= (a->a_taps[s] != 1);
    p.seg_taps[s] = a->a_taps[s];
    p.seg_cblk[s] = a->a_ch[s] / BK;
    ktot += a->a_taps[s] * a->a_ch[s];
  }
  LS_CHECK(ktot == a->Ktot, "ls_gemm: Ktot %d != sum of segments %d", a->Ktot, ktot);
  p.num_kb = ktot / BK;

  // M tile = TMA box of 128 pixels
  if (a->W >= BM) {
    p.bw = BM;
    p.bh = 1;
    p.bn = 1;
  } else {
    LS_CHECK((a->W & (a->W - 1)) == 0, "ls_gemm: W=%d (<128) must be a power of two", a->W);
    p.bw = a->W;
    int bh = BM / a->W;
    if (bh > a->H) {
      // whole images per tile: H must divide the remaining rows
      LS_CHECK((a->H & (a->H - 1)) == 0, "ls_gemm: H=%d must be a power of two when W*H < 128", a->H);
      bh = a->H;
    }
    p.bh = bh;
    p.bn = BM / (p.bw * p.bh);
  }
  p.tiles_x = (a->W + p.bw - 1) / p.bw;
  p.tiles_y = (a->H + p.bh - 1) / p.bh;
  const int tiles_n = (a->nimg + p.bn - 1) / p.bn;
  p.m_tiles = p.tiles_x * p.tiles_y * tiles_n;
  if (any_conv) LS_CHECK(a->b_batch_stride == 0, "ls_gemm: batched conv unsupported");
  p.N = a->N;
  p.b_batched = a->b_batch_stride != 0 ? 1 : 0;
  if (p.b_batched) LS_CHECK(p.bn == 1, "ls_gemm: batched GEMM needs >= 128 rows per problem or H == 1");

  // output path: smem-staged TMA bulk stores when the tile's 128 rows are contiguous output rows and the row pitch
  // is 16-byte aligned; otherwise (fp32 output, ragged geometry, narrow ld) per-thread direct stores
  const int64_t M = (int64_t)a->nimg * a->H * a->W;
  p.M = M;
  const bool rows_contig = (a->W < BM) || (a->W % BM == 0) || (a->H == 1 && a->nimg == 1);
  p.tma_store = (!(a->flags & LS_EPI_OUT_F32) && (a->ldo % 8 == 0) && rows_contig && M < (1ll << 31) &&
                 (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (a->bias_div == 0 || a->bias_div % BM == 0) &&
                 (a->residual == nullptr ||
                  ((a->ldr % 8 == 0) && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 &&
                   ((a->flags & LS_EPI_GEGLU ? a->N / 2 : a->N) % 8 == 0))))
                    ? 1
                    : 0;

  p.cs = 1;
  p.tap9_off = -1;
  if (a->stride2 != 0) {
    // 3x3 convolution with stride 2 over the [nimg, 2H, 2W] input: output pixel (y, x) reads input rows 2y - pad .. 2y - pad + 2
    LS_CHECK(a->stride2 == 1 && a->up2 == 0 && (a->stride2_pad == 0 || a->stride2_pad == 1) && a->nseg == 1 && a->a_taps[0] == 9,
             "ls_gemm: stride2 needs one 9-tap segment, stride2_pad 0 or 1, no up2");
    LS_CHECK(2 * p.bw <= 256 && 2 * p.bh <= 256, "ls_gemm: stride2 box %d x %d pixels exceeds the TMA box limit", 2 * p.bw, 2 * p.bh);
    p.cs = 2;
    p.tap9_off = -a->stride2_pad;
  }
  if (a->up2 != 0) {
    // sub-pixel phase of upsample -> conv: rows are low-resolution pixels, stored to every second pixel / row of `out`
    p.up2 = a->up2;
    p.tap_dy0 = ((a->up2 - 1) >> 1) ? 0 : -1;  // py == 0: rows y - 1, y;  py == 1: rows y, y + 1
    p.tap_dx0 = ((a->up2 - 1) & 1) ? 0 : -1;
    p.up2_bx = p.bw < 32 ? p.bw : 32;
    p.up2_by = 32 / p.up2_bx;
    LS_CHECK(a->up2 >= 1 && a->up2 <= 4 && p.tma_store && a->residual == nullptr && !(a->flags & LS_EPI_GEGLU) &&
                 p.bh % p.up2_by == 0 && (p.bw % p.up2_bx) == 0 && a->gn_partials_out == nullptr &&
                 a->row_partials_in == nullptr && a->row_partials_out == nullptr,
             "ls_gemm: up2 needs the staged fp16 epilogue without residual / GEGLU / partials and 32-pixel warp boxes "
             "(W = %d, H = %d)", a->W, a->H);
  }
  // ---- tile width, CTA pairing and split-K factor
  const int sms = num_sms();
  const bool geglu = (a->flags & LS_EPI_GEGLU) != 0;
  // folded LayerNorm (see LsGemmArgs): 1 = this GEMM consumes the producer's row partials, 2 = it produces them
  // ... 3 = it produces GroupNorm partials of its output (gn_partials_out)
  const int ln_mode = a->row_partials_in != nullptr ? 1 : (a->row_partials_out != nullptr ? 2 : (a->gn_partials_out != nullptr ? 3 : 0));
  if (ln_mode != 0) {
    LS_CHECK((a->row_partials_in != nullptr) + (a->row_partials_out != nullptr) + (a->gn_partials_out != nullptr) == 1,
             "ls_gemm: a GEMM consumes LayerNorm partials, produces them, or produces GroupNorm partials - one of the three");
    LS_CHECK(p.tma_store && a->N % 32 == 0 && !p.b_batched,
             "ls_gemm: LayerNorm partials need the staged fp16 epilogue (ldo %% 8 == 0, contiguous rows), N %% 32 == 0, no batching");
  }
  if (ln_mode == 1) {
    LS_CHECK(a->residual == nullptr, "ls_gemm: a GEMM with a folded LayerNorm takes no residual");
    LS_CHECK((reinterpret_cast<uintptr_t>(a->col_sum) & 15) == 0, "ls_gemm: col_sum must be 16-byte aligned");
    LS_CHECK(a->col_sum != nullptr && a->n_partials_in >= 1 && a->n_partials_in <= 96 && a->partials_in_stride >= M &&
                 a->ln_eps > 0.f && a->nseg == 1 && a->a_taps[0] == 1 &&
                 (reinterpret_cast<uintptr_t>(a->row_partials_in) & 7) == 0,
             "ls_gemm: row_partials_in needs col_sum, 1..96 parts with stride >= M, ln_eps > 0 and a single pointwise A segment");
  }
  if (ln_mode == 2) {
    LS_CHECK(!geglu && !(a->flags & LS_EPI_SILU) && a->partials_out_stride >= M &&
                 (reinterpret_cast<uintptr_t>(a->row_partials_out) & 7) == 0,
             "ls_gemm: row_partials_out needs a plain (bias / residual) epilogue and stride >= M");
  }
  if (ln_mode == 3) {
    LS_CHECK(!geglu && !(a->flags & LS_EPI_SILU) && M % BM == 0 && a->gn_unit >= 1 && a->N % a->gn_unit == 0 &&
                 a->gn_partials_ld >= a->N / a->gn_unit && (reinterpret_cast<uintptr_t>(a->gn_partials_out) & 7) == 0,
             "ls_gemm: gn_partials_out needs a plain (bias / residual) epilogue, M %% 128 == 0, N %% gn_unit == 0, "
             "gn_partials_ld >= N / gn_unit");
  }
  static int env_ctas = -1;
  if (env_ctas < 0) {
    const char* e = getenv("LS_GEMM_CTAS");
    env_ctas = e ? atoi(e) : 0;
  }
  const int force_ctas = a->cta_pair != 0 ? a->cta_pair : env_ctas;
  // a pair shares one B tile: both 128-row halves must belong to the same batched problem
  const bool pair_ok = (p.m_tiles >= 2) && (!p.b_batched || ((p.tiles_x * p.tiles_y) % 2 == 0)) && (sms % 2 == 0);
  static int env_split = -1;
  if (env_split < 0) {
    const char* e = getenv("LS_GEMM_SPLITK");  // 0 disables split-K (A/B measurements)
    env_split = e ? atoi(e) : 1;
  }
  // split-K only where SMs would idle: few output tiles, long K; needs the TMA-store epilogue and a plain GEMM
  const bool split_ok = env_split != 0 && p.tma_store && !geglu && !p.b_batched && !(a->flags & LS_EPI_SILU) &&
                        (a->N % 8 == 0) && (a->bias_ld % 4 == 0) && ln_mode == 0 && a->up2 == 0;
  int best_bn = 0, best_ctas = 1, best_split = 1;
  double best_cost = -1.0;
  const int step = geglu ? 64 : 32;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (ctas == 2 && !pair_ok) continue;
    if (force_ctas == 1 && ctas == 2) continue;
    if (force_ctas == 2 && ctas == 1 && pair_ok) continue;
    for (int bn = step; bn <= 256; bn += step) {
      if (a->tile_n != 0 && bn != a->tile_n) continue;
      if (geglu && (a->N % bn != 0)) continue;
      if (bn > 32 && bn - 32 >= a->N && a->tile_n == 0) continue;  // wider than the problem
      // GroupNorm partials: whole units per tile, and no tile cut short by the end of N
      if (ln_mode == 3 && (a->N % bn != 0 || bn % a->gn_unit != 0)) continue;
      for (int sp = 1; sp <= 8; ++sp) {
        if (env_split >= 2 && split_ok && ctas == 1 && sp != env_split && p.num_kb >= 4 * env_split) continue;  // forced
        if (sp > 1) {
          if (!split_ok || ctas != 1) break;
          const int kb_per = (p.num_kb + sp - 1) / sp;
          if (kb_per < 4 || (sp - 1) * kb_per >= p.num_kb) continue;        // every split needs work
          if ((double)sp * M * a->N * 4.0 > 64.0 * 1024 * 1024) break;      // workspace bound
          const long tiles = (long)p.m_tiles * ((a->N + bn - 1) / bn);
          if (tiles * sp > sms) break;  // the rendezvous needs every CTA of the launch resident: a single wave
        }
        const double c = tile_cost(ctas, bn, sp, p.m_tiles, a->N, p.num_kb, sms, geglu, a->residual != nullptr);
        if (best_cost < 0 || c < best_cost) {
          best_cost = c;
          best_bn = bn;
          best_ctas = ctas;
          best_split = sp;
        }
      }
    }
  }
  LS_CHECK(best_bn > 0, "ls_gemm: no valid tile width for N=%d tile_n=%d flags=%d", a->N, a->tile_n, a->flags);
  const int BN = best_bn, CTAS = best_ctas;
  p.BN = BN;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.splits = best_split;
  p.kb_per = (p.num_kb + best_split - 1) / best_split;
  if (ln_mode == 2)
    LS_CHECK(a->n_partials_out == 3 * p.n_tiles,
             "ls_gemm: row_partials_out holds n_partials_out = %d parts, this launch writes 3 * ceil(N / tile_n) = %d (N = %d, "
             "tile_n = %d): pass tile_n explicitly", a->n_partials_out, 3 * p.n_tiles, a->N, BN);
  p.ln_in = reinterpret_cast<const float2*>(a->row_partials_in);
  p.colsum = a->col_sum;
  p.ln_nparts_in = a->n_partials_in;
  p.ln_in_stride = a->partials_in_stride;
  p.ln_eps = a->ln_eps;
  p.ln_inv_k = 1.0f / (float)ktot;
  p.ln_out = reinterpret_cast<float2*>(a->row_partials_out);
  p.ln_out_stride = a->partials_out_stride;
  if (ln_mode == 3) {
    p.gn_out = reinterpret_cast<float2*>(a->gn_partials_out);
    p.gn_unit = a->gn_unit;
    p.gn_ld = a->gn_partials_ld;
    p.gn_upt = BN / a->gn_unit;
  }
  if (best_split > 1) {
    int dev = 0;
    LS_CUDA(cudaGetDevice(&dev));
    LS_CHECK(dev >= 0 && dev < 16, "ls_gemm: device index %d out of range", dev);
    SplitScratch& sc = g_split[dev];
    const size_t need = (size_t)best_split * M * a->N;
    const size_t ntick = (size_t)p.m_tiles * p.n_tiles * 2 + 2;
    // 64 MB of fp32 partials / 8 K tickets cover every UNet / VAE shape; a superseded block is never freed
    int rc = scratch_reserve(sc.ws_b, need * sizeof(float), (size_t)64 << 20, false, stream, "ls_gemm");
    if (rc != 0) return rc;
    rc = scratch_reserve(sc.tickets_b, ntick * sizeof(unsigned int), (size_t)8192 * sizeof(unsigned int), true, stream,
                         "ls_gemm");
    if (rc != 0) return rc;
    p.ws = reinterpret_cast<float*>(sc.ws_b.ptr);
    p.tickets = reinterpret_cast<unsigned int*>(sc.tickets_b.ptr);
  }

  // tensor maps
  for (int s = 0; s < a->nseg; ++s) {
    const cuuint64_t ld_b = (cuuint64_t)a->a_ld[s] * 2;
    // stride2: the map covers the INPUT image (2H x 2W); the box spans 2 bw x 2 bh input pixels of which every second one
    // (element strides 2) is loaded, i.e. bw x bh pixels land in shared memory as usual
    const cuuint64_t cs = (cuuint64_t)p.cs;
    cuuint64_t gdim[4] = {(cuuint64_t)a->a_ch[s], cs * a->W, cs * a->H, (cuuint64_t)a->nimg};
    cuuint64_t gstr[3] = {ld_b, ld_b * cs * a->W, ld_b * cs * a->W * cs * a->H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(p.cs * p.bw), (cuuint32_t)(p.cs * p.bh), (cuuint32_t)p.bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)p.cs, (cuuint32_t)p.cs, 1};
    CUresult r = encode(&p.mapA[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(a->a_ptr[s]), gdim, gstr,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(A%d) failed with %d", s, (int)r);
  }
  {
    const cuuint64_t nb = p.b_batched ? (cuuint64_t)a->nimg : 1;
    const cuuint64_t bstride = p.b_batched ? (cuuint64_t)a->b_batch_stride * 2 : (cuuint64_t)a->N * ktot * 2;
    LS_CHECK((reinterpret_cast<uintptr_t>(a->b_ptr) & 15) == 0 && (bstride & 15) == 0, "ls_gemm: B alignment");
    cuuint64_t gdim[3] = {(cuuint64_t)ktot, (cuuint64_t)a->N, nb};
    cuuint64_t gstr[2] = {(cuuint64_t)ktot * 2, bstride};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BN / CTAS), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(a->b_ptr), gdim, gstr, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }

  const int n_out = geglu ? a->N / 2 : a->N;
  if (p.tma_store && p.up2) {
    // [column][x][y][image] view of the HIGH-resolution output (2H x 2W pixels per image, row pitch ldo), every second
    // pixel and row, starting at this phase's pixel (py, px); one warp stores a box of up2_bx x up2_by pixels
    const int py = (a->up2 - 1) >> 1, px = (a->up2 - 1) & 1;
    const cuuint64_t ldb = (cuuint64_t)a->ldo * 2;
    __half* base = reinterpret_cast<__half*>(a->out) + ((int64_t)py * (2 * a->W) + px) * a->ldo;
    LS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "ls_gemm: up2 output alignment");
    cuuint64_t gdim[4] = {(cuuint64_t)a->N, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->nimg};
    cuuint64_t gstr[3] = {2 * ldb, 2 * (cuuint64_t)(2 * a->W) * ldb, (cuuint64_t)(2 * a->H) * (2 * a->W) * ldb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint32_t box[4] = {64, (cuuint32_t)p.up2_bx, (cuuint32_t)p.up2_by, 1};
    CUresult r = encode(&p.mapOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(out, up2) failed with %d", (int)r);
    cuuint32_t box32[4] = {32, (cuuint32_t)p.up2_bx, (cuuint32_t)p.up2_by, 1};
    r = encode(&p.mapOut32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, gdim, gstr, box32, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(out32, up2) failed with %d", (int)r);
  } else if (p.tma_store) {
    const int n_out = geglu ? a->N / 2 : a->N;
    cuuint64_t gdim[2] = {(cuuint64_t)n_out, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)a->ldo * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.mapOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->out, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    cuuint32_t box32[2] = {32, 32};
    r = encode(&p.mapOut32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->out, gdim, gstr, box32, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(out32) failed with %d", (int)r);
  }
  p.bias = a->bias;
  p.bias_div = a->bias_div;
  p.fd_tiles_x = make_fastdiv((uint32_t)p.tiles_x);
  p.fd_tiles_y = make_fastdiv((uint32_t)p.tiles_y);
  p.fd_n_tiles = make_fastdiv((uint32_t)p.n_tiles);
  p.fd_splits = make_fastdiv((uint32_t)p.splits);
  p.fd_bias_div = make_fastdiv((uint32_t)(a->bias_div > 0 ? a->bias_div : 1));
  p.bias_ld = a->bias_ld > 0 ? a->bias_ld : a->N;
  p.residual = reinterpret_cast<const __half*>(a->residual);
  p.ldr = a->ldr;
  p.out = a->out;
  p.ldo = a->ldo;
  p.flags = a->flags;

  static int env_kbs = -1;
  if (env_kbs < 0) {
    // 1 / 2 forces the k-blocks per pipeline stage.  Default 1: two per stage are 3-12 % SLOWER on every UNet / VAE shape
    // (profiles/r2e_gemm_kbs.txt) - the coarser refill granularity costs more than the halved hand-shake saves.
    const char* e = getenv("LS_GEMM_KBS");
    env_kbs = e ? atoi(e) : 0;
  }
  // two k-blocks per stage halve the per-stage mbarrier hand-shake; they need at least two stages' worth of work
  p.kbs = (env_kbs == 1 || env_kbs == 2) ? env_kbs : LS_GEMM_DEFAULT_KBS;
  if (p.kb_per < 4) p.kbs = 1;
  const int stage_bytes = p.kbs * (A_STAGE_BYTES + (BN / CTAS) * 128);
  int fixed = STAGING_BYTES + BIAS_BYTES + 192 + (ln_mode == 1 ? COLSUM_BYTES : 0) +
              (ln_mode == 3 ? GN_SM_BYTES_PER_COL * BN : 0);  // + 21 mbarriers and the TMEM slot
  int stages = (SMEM_BUDGET - fixed) / stage_bytes;
  // Weight-resident mode for short-K launches whose whole B tile fits beside an A ring (K = 320 at BN = 160: 100 KB): the
  // grid is trimmed to a multiple of n_tiles so that CTA c only ever sees n-tile c % n_tiles.  LS_GEMM_BRES=0 disables.
  static int env_bres = -1;
  if (env_bres < 0) {
    const char* e = getenv("LS_GEMM_BRES");
    env_bres = e ? atoi(e) : LS_GEMM_DEFAULT_BRES;
  }
  int bres_bytes = 0;
  {
    const int units0 = sms;
    const long total0 = (long)p.m_tiles * p.n_tiles;
    const int g_units = (units0 / p.n_tiles) * p.n_tiles;
    const int b_bytes = p.num_kb * BN * 128;
    const int a_stages = (SMEM_BUDGET - fixed - b_bytes) / A_STAGE_BYTES;
    if (env_bres != 0 && ln_mode == 0 && CTAS == 1 && p.splits == 1 && p.kbs == 1 && !p.b_batched && g_units > 0 &&
        b_bytes <= 112 * 1024 && a_stages >= 4 && total0 >= 2L * g_units) {
      p.bres = 1;
      bres_bytes = b_bytes;
      stages = a_stages;
    }
  }
  // Experiment kept behind LS_GEMM_SLAB_SINGLE: if main-loop-bound launches were bound by the operand bytes in flight,
  // halving the staging area (slabs stored one by one from 2 KB per warp) to buy one more stage would pay.  It does not.
  static int env_ss = -1;
  if (env_ss < 0) {
    // 1 forces it where legal, 2 = where it buys a stage.  Default OFF: measured neutral (profiles/r2f_gemm_slab_single.txt:
    // 7 -> 8 stages at 256 x 160 changes the 3x3 convolutions by < 1 %), i.e. the main loop is not bound by bytes in flight.
    const char* e = getenv("LS_GEMM_SLAB_SINGLE");
    env_ss = e ? atoi(e) : 0;
  }
  {
    const int fixed_s = STAGING_BYTES / 2 + BIAS_BYTES + 192;
    const int stages_s = (SMEM_BUDGET - fixed_s) / stage_bytes;
    const bool legal = p.tma_store && a->residual == nullptr && p.splits == 1 && !geglu && !p.bres && ln_mode == 0;
    const bool want = env_ss == 1 || (env_ss == 2 && p.kb_per >= 24 && stages_s > stages && stages < MAX_STAGES);
    if (legal && want) {
      p.slab_single = 1;
      fixed = fixed_s;
      stages = stages_s;
    }
  }
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = p.bres ? (size_t)bres_bytes + (size_t)stages * A_STAGE_BYTES + fixed : (size_t)stages * stage_bytes + fixed;

  const int m_units = (p.m_tiles + CTAS - 1) / CTAS;
  const long total = (long)m_units * p.n_tiles * p.splits;
  const int units = sms / CTAS;
  static int env_alt = -1;
  if (env_alt < 0) {
    const char* e = getenv("LS_GEMM_EPI_ALT");  // 0 / 1 forces the epilogue mode (A/B measurements)
    env_alt = e ? atoi(e) : 2;
  }
  p.epi_alt = (p.tma_store && p.splits == 1 && (env_alt == 1 || (env_alt == 2 && total >= 3L * units))) ? 1 : 0;
  int grid = (int)(total < units ? total : units) * CTAS;
  if (p.bres) grid = (units / p.n_tiles) * p.n_tiles;  // every tile of a CTA shares its n-tile (total >= 2 x that, checked above)
  static bool attr_set_dev[16] = {};  // function attributes are per device
  int dev_attr = 0;
  LS_CUDA(cudaGetDevice(&dev_attr));
  LS_CHECK(dev_attr >= 0 && dev_attr < 16, "ls_gemm: device index %d out of range", dev_attr);
  bool& attr_set = attr_set_dev[dev_attr];
  if (!attr_set) {
    for (int c = 1; c <= 2; ++c)
      for (int g = 0; g < 2; ++g)
        for (int l = 0; l < 4; ++l)
          if (GemmKernelFn fn = gemm_kernel_of(c, g != 0, l))
            LS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    attr_set = true;
  }
  if (CTAS == 1 && p.splits > 1)
    // split-K CTAs meet at a per-tile ticket inside the kernel: cooperative launch = the driver guarantees that the
    // whole (single-wave) grid is resident, whatever else runs on the GPU
    LS_CUDA(launch_coop_k(gemm_tc_kernel<false, 0>, dim3(grid), dim3(GEMM_THREADS), (size_t)(smem), (cudaStream_t)(stream),
                          p));
  else
    LS_CUDA(launch_k(gemm_kernel_of(CTAS, geglu, ln_mode), dim3(grid), dim3(GEMM_THREADS), (size_t)(smem),
                     (cudaStream_t)(stream), p));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace ls

extern "C" int ls_gemm(const LsGemmArgs* args, void* stream) {
  return ls::gemm_impl(args, reinterpret_cast<cudaStream_t>(stream));
}
