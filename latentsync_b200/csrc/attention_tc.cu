// Flash-style attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the long-sequence spatial
// self-attention of the UNet (attention.py:271: S = H*W = 1024 / 256 / 4096 tokens per frame, head_dim 40 / 80 / 160).
//
//   per CTA: one 128-query tile of one (frame, head); loop over 64-key tiles
//     control warp (1 elected lane): TMA loads of Q / K_j / V_j (4-D maps over [batch][seq][head][d]: the box is 64
//                                    columns wide, columns >= head_dim and rows >= seq are zero-filled by TMA),
//                                    S_j = Q K_j^T (tcgen05.mma, both operands K-major, 128B swizzle) -> TMEM,
//                                    O  += P_j V_j (A = P_j from shared memory, B = V_j MN-major: V is used as it lies
//                                    in memory, [key][d], no transpose) -> TMEM
//     4 softmax warps (thread = query row = TMEM lane): tcgen05.ld S_j -> registers, release S early so that
//                                    S_{j+1} overlaps the exponentials, running max with LAZY rescale (the reference
//                                    max only moves when it grows by more than 2^8; O in TMEM is then rescaled with
//                                    tcgen05.ld/st), P_j = exp2(...) as fp16 into the swizzled shared-memory A tile.
//   The exponentials (MUFU, 16 per clock per SM) bound this kernel, not the tensor pipe: 128 x 64 exp per tile = 512
//   clocks against 192 clocks of MMA (head_dim 40), so several CTAs are resident per SM to keep the MUFU pipe busy.
//
// Replaces F.scaled_dot_product_attention at latentsync/models/attention.py:271 and motion_module.py:300:
//   * spatial self-attention and the audio cross-attention (50 keys: one key tile whose columns >= skv are masked) at
//     every level - a query tile with fewer than 128 valid rows (8x8 / 4x4 levels) simply leaves TMEM lanes unused;
//   * PACK = true: the temporal attention over F = 16 frames.  A 128-row tile holds 128 / F PIXELS x F frames, gathered by
//     ONE 5-D TMA box over the [(b f) (h w)] x C token matrix (frame stride = h*w rows, pixel stride = 1 row: the
//     "(b f) s c -> (b s) f c" rearrange of motion_module.py:264 is never materialised), rows ordered pixel-major so that
//     S = Q K^T is block diagonal with F x F blocks; the other blocks are masked before the softmax.  8 sequences share
//     one pair of tensor-core GEMMs instead of one warp-level mma.sync problem each.
// The warp-level kernels of attention.cu remain only as the fallback for shapes this path does not take.
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace ls {

extern std::atomic<int64_t> g_launch_count;

struct AttnTcParams {
  CUtensorMap mapQ, mapK, mapV;
  __half* o;
  int ldo;
  int sq, skv;
  int64_t o_batch_stride;  // rows
  int64_t o_seq_stride;    // rows
  float scale_log2;
  int pack_flog;   // PACK: log2(frames per sequence)
  int pack_hw;     // PACK: pixels per batch element (extent of the pixel dimension)
  long long* probe;  // LS_ATC_PROBE builds only: clock64 stamps [cta][tile][8]
};

// masked score in PACK mode.  Finite on purpose: a row whose block lies in the SECOND key tile sees a fully masked first
// tile; with -inf its running maximum would stay -inf and exp2(-inf - -inf) = NaN.  With a finite value the first tile
// yields P = 1 garbage that the rescale of the next tile multiplies by exp2((-3e4 - m) * scale) = 0.  Small enough that
// the rounding error of NEG * scale_log2 stays far below 1 (so exp2(~0) = 1, not inf).
constexpr float ATC_PACK_NEG = -30000.0f;

#ifdef LS_ATC_PROBE
#define ATC_STAMP(slot) \
  do { if (p.probe && blockIdx.y == 0 && (blockIdx.z == 0 || blockIdx.z == 16)) \
      p.probe[(((blockIdx.z ? 8 : 0) + blockIdx.x) * 64 + j) * 16 + (slot)] = clock64(); } while (0)
#else
#define ATC_STAMP(slot) do {} while (0)
#endif

constexpr int ATC_BQ = 128;
constexpr int ATC_BKV = 64;
constexpr int ATC_THREADS = 160;  // 4 softmax warps + 1 control warp
constexpr float ATC_RESCALE_THRESHOLD = 8.0f;  // in log2 units: P <= 2^8 stays far inside fp16 range

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand, 128-byte swizzle: the tile lies in shared memory as [k][64 MN elements] (128-byte rows, 8-row
// groups 1024 B apart = stride byte offset); 64-element MN blocks are `lbo_bytes` apart (leading byte offset).
// Canonical layout Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int D>
struct AtcCfg {
  static constexpr int DP = (D + 15) / 16 * 16;        // head_dim padded to the MMA K step / N granularity
  static constexpr int KS = (DP + 63) / 64;            // 64-column slabs of Q / K / V
  static constexpr int KSTEPS = DP / 16;               // K = 16 steps of S = Q K^T
  static constexpr int Q_BYTES = KS * ATC_BQ * 128;
  static constexpr int KV_SLAB = ATC_BKV * 128;
  static constexpr int KV_BYTES = KS * KV_SLAB;        // one stage of K (or V)
  static constexpr int P_BYTES = ATC_BQ * 128;         // 128 rows x 64 keys fp16
  static constexpr int TILE_BYTES = Q_BYTES + 4 * KV_BYTES + P_BYTES;
  static constexpr int SMEM = TILE_BYTES + 1024 /*alignment*/ + 128 /*barriers*/;
  static constexpr int TMEM_NEED = ATC_BKV + DP;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int CTAS_PER_SM = (D <= 40) ? 3 : (D <= 80 ? 2 : 1);
};

template <int D, bool PACK>
__global__ void __launch_bounds__(ATC_THREADS, AtcCfg<D>::CTAS_PER_SM) attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  using Cfg = AtcCfg<D>;
  constexpr int DP = Cfg::DP, KS = Cfg::KS, KSTEPS = Cfg::KSTEPS;
  pdl_prologue();
  extern __shared__ uint8_t atc_smem_raw[];
  const uint32_t raw = smem_u32(atc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = atc_smem_raw + (base - raw);
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + 2 * Cfg::KV_BYTES;
  const uint32_t sP = sV + 2 * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // PACK: the tile's 128 keys are its own 128 rows (pixel-major): two key tiles of 64 = 64 / F pixels each
  const int n_tiles = PACK ? ATC_BQ / ATC_BKV : (p.skv + ATC_BKV - 1) / ATC_BKV;
  const int ppt = PACK ? (ATC_BQ >> p.pack_flog) : 0;      // pixels per query tile
  const int ppk = PACK ? (ATC_BKV >> p.pack_flog) : 0;     // pixels per key tile

  if (tid == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full + 0, 1);
    mbar_init(k_full + 1, 1);
    mbar_init(v_full + 0, 1);
    mbar_init(v_full + 1, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem;              // columns [0, 64): scores
  const uint32_t tO = tmem + ATC_BKV;    // columns [64, 64 + DP): output accumulator

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------------------------------------ control: TMA producer + MMA issuer
      tma_prefetch_desc(&p.mapQ);
      tma_prefetch_desc(&p.mapK);
      tma_prefetch_desc(&p.mapV);
      constexpr uint32_t idesc_s = (1u << 4) | (uint32_t(ATC_BKV >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 16) | (uint32_t(DP >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      auto load_k = [&](int j, int st) {
        mbar_expect_tx(k_full + st, Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if constexpr (PACK)
            tma_load_5d(sm + (sK - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapK, k_full + st, s * 64, h, 0,
                        qt * ppt + j * ppk, b);
          else
            tma_load_4d(sm + (sK - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapK, k_full + st, s * 64, h,
                        j * ATC_BKV, b);
        }
      };
      auto load_v = [&](int j, int st) {
        mbar_expect_tx(v_full + st, Cfg::KV_BYTES);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if constexpr (PACK)
            tma_load_5d(sm + (sV - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapV, v_full + st, s * 64, h, 0,
                        qt * ppt + j * ppk, b);
          else
            tma_load_4d(sm + (sV - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapV, v_full + st, s * 64, h,
                        j * ATC_BKV, b);
        }
      };
      auto issue_s = [&](int st) {
        // S = Q K^T: KSTEPS steps of K = 16; step ks lies in slab ks / 4 at byte offset (ks % 4) * 32
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t qa = sQ + (ks >> 2) * (ATC_BQ * 128) + (ks & 3) * 32;
          const uint32_t ka = sK + st * Cfg::KV_BYTES + (ks >> 2) * Cfg::KV_SLAB + (ks & 3) * 32;
          umma_f16_ss(tS, umma_desc_sw128(qa), umma_desc_sw128(ka), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      auto issue_pv = [&](int st, uint32_t accumulate) {
        // O += P V: 4 steps of 16 keys; A = P [128][64 keys] K-major, B = V [64 keys][DP] MN-major
#pragma unroll
        for (int kk = 0; kk < ATC_BKV / 16; ++kk) {
          const uint32_t pa = sP + kk * 32;
          const uint32_t va = sV + st * Cfg::KV_BYTES + kk * 16 * 128;
          umma_f16_ss(tO, umma_desc_sw128(pa), umma_desc_mn_sw128(va, Cfg::KV_SLAB), idesc_o,
                      (accumulate | (uint32_t)(kk > 0)) ? 1u : 0u);
        }
        umma_commit(pv_done);
      };

      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        if constexpr (PACK)
          tma_load_5d(sm + (sQ - base) + s * (ATC_BQ * 128), &p.mapQ, q_full, s * 64, h, 0, qt * ppt, b);
        else
          tma_load_4d(sm + (sQ - base) + s * (ATC_BQ * 128), &p.mapQ, q_full, s * 64, h, qt * ATC_BQ, b);
      }
      load_k(0, 0);
      load_v(0, 0);
      if (n_tiles > 1) load_k(1, 1);
      mbar_wait(q_full, 0);
      mbar_wait(k_full + 0, 0);
      tc_fence_after();
      issue_s(0);
      // Per key tile: 4 barrier waits.  Completion of earlier tensor work is inferred instead of waited for:
      //   s_empty(j) (the softmax warps hold S_j)      => S_j is complete      => K stage j&1 can be refilled
      //   p_full(j)  (they waited for P V_{j-1} first) => P V_{j-1} is complete => V stage (j+1)&1 can be refilled
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)(j & 1);
        ATC_STAMP(8);
        if (j + 1 < n_tiles) {
          mbar_wait(s_empty, ph);
          ATC_STAMP(9);
          if (j + 2 < n_tiles) load_k(j + 2, st);
          ATC_STAMP(10);
          mbar_wait(k_full + (st ^ 1), (uint32_t)(((j + 1) >> 1) & 1));
          tc_fence_after();
          ATC_STAMP(0);
          issue_s(st ^ 1);
          ATC_STAMP(12);
        }
        mbar_wait(p_full, ph);
        ATC_STAMP(11);
        if (j + 1 < n_tiles) load_v(j + 1, st ^ 1);
        mbar_wait(v_full + st, (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        ATC_STAMP(1);
        issue_pv(st, j > 0 ? 1u : 0u);
        ATC_STAMP(14);
      }
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------------- softmax: thread = query row
    const int r = tid;  // 0..127 = TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl = p.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    const uint32_t p_row = sP + r * 128;
    const int rx = r & 7;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      if (tid == 0) ATC_STAMP(3);
      mbar_wait(s_full, ph);
      tc_fence_after();
      if (tid == 0) ATC_STAMP(4);
      uint32_t sv[2][32];
      tmem_ld_32x32(tS + lane_off, sv[0]);
      tmem_ld_32x32(tS + lane_off + 32, sv[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      if (tid == 0) ATC_STAMP(5);
      float* s = reinterpret_cast<float*>(&sv[0][0]);
      const int kbase = j * ATC_BKV;
      if constexpr (PACK) {
        // block-diagonal mask: row r = pixel (r >> flog) attends to the keys of the same pixel only
        const int my_blk = r >> p.pack_flog;
#pragma unroll
        for (int c = 0; c < ATC_BKV; ++c)
          if (((kbase + c) >> p.pack_flog) != my_blk) s[c] = ATC_PACK_NEG;
      } else if (kbase + ATC_BKV > p.skv) {
#pragma unroll
        for (int c = 0; c < ATC_BKV; ++c)
          if (kbase + c >= p.skv) s[c] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < ATC_BKV; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const bool grow = (mx - m_ref) * sl > ATC_RESCALE_THRESHOLD;  // first tile: m_ref = -inf -> true
      bool waited_pv = false;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_ref, mx);
        const float alpha = ex2_approx((m_ref - m_new) * sl);  // 0 on the first tile
        m_ref = m_new;
        l *= alpha;
        if (j > 0) {
          mbar_wait(pv_done, (uint32_t)((j - 1) & 1));  // O holds tiles 0..j-1
          tc_fence_after();
          waited_pv = true;
#pragma unroll
          for (int c = 0; c < DP; c += 16) {
            uint32_t ov[16];
            tmem_ld_32x16(tO + lane_off + c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(tO + lane_off + c, ov);
          }
          tmem_st_wait();
        }
      }
      const float mb = m_ref * sl;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[ATC_BKV / 2];
#pragma unroll
      for (int c = 0; c < ATC_BKV; c += 2) {
        const float e0 = ex2_approx(fmaf(s[c], sl, -mb));
        const float e1 = ex2_approx(fmaf(s[c + 1], sl, -mb));
        sum0 += e0;
        sum1 += e1;
        pk[c >> 1] = pack_half2(e0, e1);
      }
      l += sum0 + sum1;
      if (tid == 0) ATC_STAMP(6);
      if (j > 0 && !waited_pv) mbar_wait(pv_done, (uint32_t)((j - 1) & 1));  // P V_{j-1} has read the P tile
#pragma unroll
      for (int c = 0; c < ATC_BKV / 8; ++c)
        st_shared_v4(p_row + ((c ^ rx) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tid == 0) ATC_STAMP(7);
    }
    // epilogue: O / l -> fp16 -> global (each thread owns one output row of D contiguous halfs)
    mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    int qrow = qt * ATC_BQ + r;
    __half* dst;
    if constexpr (PACK) {
      // row r = (pixel r >> flog, frame r & (F - 1)); o_batch_stride = rows per batch element, o_seq_stride = rows per
      // frame, pixels are consecutive rows
      const int pix = qt * ppt + (r >> p.pack_flog);
      const int fr = r & ((1 << p.pack_flog) - 1);
      dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)fr * p.o_seq_stride + pix) * p.ldo + h * D;
      qrow = (pix < p.pack_hw) ? 0 : p.sq;  // the guard below: rows of pixels past the image are not stored
    } else {
      dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)qrow * p.o_seq_stride) * p.ldo + h * D;
    }
#pragma unroll
    for (int c = 0; c < DP; c += 16) {
      uint32_t ov[16];
      tmem_ld_32x16(tO + lane_off + c, ov);
      tmem_ld_wait();
      if (qrow < p.sq) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < D) {
            uint4 pk4;
            pk4.x = pack_half2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
            pk4.y = pack_half2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            pk4.z = pack_half2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
            pk4.w = pack_half2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + i) = pk4;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}

// =================================================================================================================
// Version 2: everything the softmax warps wait for is produced a whole key tile ahead.
//   * S is double buffered in TMEM (S_{j+2} is issued when the warps have finished tile j, so S_{j+1} is already there
//     when they come back), P is double buffered in shared memory (no wait for P V_{j-1} before writing P_j),
//     K / V live in rings of NST stages loaded two tiles ahead.
//   * ONE barrier per key tile, R[j & 1]: 4 warp arrivals ("P_j is written" - which implies "S_j has been read") plus
//     the transaction bytes of the TMA loads that the tensor work issued at that wake-up needs (V_j, K_{j+2}).  The
//     control thread wakes once per tile and issues P V_j, S_{j+2} and the loads for tile j + 2.  Version 1 needed
//     4-6 waits per tile, each ~300 clocks behind the MUFU backlog of the MIO queue (profiles/r1b_attn_tc_probe.txt).
//   * Completion of older tensor work is inferred, never waited for: s_full(j) is a tcgen05.commit issued after
//     P V_{j-2}, so a warp that holds S_j knows that P buffer j & 1 and V stage (j - 2) % NST are free.
// =================================================================================================================
template <int D>
struct Atc2Cfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int KS = (DP + 63) / 64;
  static constexpr int KSTEPS = DP / 16;
  static constexpr int NST = (D <= 80) ? 4 : 3;        // K / V ring depth
  static constexpr int Q_BYTES = KS * ATC_BQ * 128;
  static constexpr int KV_SLAB = ATC_BKV * 128;
  static constexpr int KV_BYTES = KS * KV_SLAB;
  static constexpr int P_BYTES = ATC_BQ * 128;
  static constexpr int TILE_BYTES = Q_BYTES + 2 * NST * KV_BYTES + 2 * P_BYTES;
  static constexpr int SMEM = TILE_BYTES + 128;        // barriers; the dynamic segment is 1024-byte aligned (checked)
  static constexpr int TMEM_NEED = 2 * ATC_BKV + DP;
  static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int CTAS_PER_SM = (2 * (SMEM + 1024) <= 228 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
};

template <int D>
__global__ void __launch_bounds__(ATC_THREADS, Atc2Cfg<D>::CTAS_PER_SM) attn_tc2_kernel(const __grid_constant__ AttnTcParams p) {
  using Cfg = Atc2Cfg<D>;
  constexpr bool PACK = false;  // version 2 only takes the long, contiguous self-attention sequences
  constexpr int DP = Cfg::DP, KS = Cfg::KS, KSTEPS = Cfg::KSTEPS, NST = Cfg::NST;
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t atc2_smem[];
  uint8_t* sm = atc2_smem;
  const uint32_t base = smem_u32(sm);
  if ((base & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("latentsync_b200: attention smem not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t sQ = base;
  const uint32_t sK = sQ + Cfg::Q_BYTES;
  const uint32_t sV = sK + NST * Cfg::KV_BYTES;
  const uint32_t sP = sV + NST * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* rdy = bars + 1;      // [2]  R[j & 1]
  uint64_t* s_full = bars + 3;   // [2]
  uint64_t* pv_done = bars + 5;
  uint64_t* fin = bars + 6;      // committed once, after the last P V
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  // PACK: the tile's 128 keys are its own 128 rows (pixel-major): two key tiles of 64 = 64 / F pixels each
  const int n_tiles = PACK ? ATC_BQ / ATC_BKV : (p.skv + ATC_BKV - 1) / ATC_BKV;
  const int ppt = PACK ? (ATC_BQ >> p.pack_flog) : 0;      // pixels per query tile
  const int ppk = PACK ? (ATC_BKV >> p.pack_flog) : 0;     // pixels per key tile

  if (tid == 0) {
    mbar_init(q_full, 1);
    mbar_init(rdy + 0, 5);
    mbar_init(rdy + 1, 5);
    mbar_init(s_full + 0, 1);
    mbar_init(s_full + 1, 1);
    mbar_init(pv_done, 1);
    mbar_init(fin, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = tmem + 2 * ATC_BKV;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&p.mapQ);
      tma_prefetch_desc(&p.mapK);
      tma_prefetch_desc(&p.mapV);
      constexpr uint32_t idesc_s = (1u << 4) | (uint32_t(ATC_BKV >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 16) | (uint32_t(DP >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      auto load_k = [&](int j, uint64_t* bar) {
        const int st = j % NST;
#pragma unroll
        for (int s = 0; s < KS; ++s)
          tma_load_4d(sm + (sK - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapK, bar, s * 64, h, j * ATC_BKV, b);
      };
      auto load_v = [&](int j, uint64_t* bar) {
        const int st = j % NST;
#pragma unroll
        for (int s = 0; s < KS; ++s)
          tma_load_4d(sm + (sV - base) + st * Cfg::KV_BYTES + s * Cfg::KV_SLAB, &p.mapV, bar, s * 64, h, j * ATC_BKV, b);
      };
      auto issue_s = [&](int j) {  // S_j = Q K_j^T into S buffer j & 1
        const uint32_t tS = tmem + (uint32_t)(j & 1) * ATC_BKV;
        const uint32_t kb = sK + (j % NST) * Cfg::KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t qa = sQ + (ks >> 2) * (ATC_BQ * 128) + (ks & 3) * 32;
          const uint32_t ka = kb + (ks >> 2) * Cfg::KV_SLAB + (ks & 3) * 32;
          umma_f16_ss(tS, umma_desc_sw128(qa), umma_desc_sw128(ka), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full + (j & 1));
      };
      auto issue_pv = [&](int j) {  // O += P_j V_j
        const uint32_t pb = sP + (uint32_t)(j & 1) * Cfg::P_BYTES;
        const uint32_t vb = sV + (j % NST) * Cfg::KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < ATC_BKV / 16; ++kk)
          umma_f16_ss(tO, umma_desc_sw128(pb + kk * 32), umma_desc_mn_sw128(vb + kk * 16 * 128, Cfg::KV_SLAB), idesc_o,
                      (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(pv_done);
      };
      // Barrier accounting: the phase of R[t & 1] that belongs to tile t gets 4 warp arrivals and ONE arrival from this
      // thread, posted two tiles earlier (prologue for t < 2) together with the bytes of V_t and K_{t+2} - the loads
      // the wake-up of tile t depends on.  With NST = 3 the K_{t+2} load itself is issued one wake-up later than its
      // bytes are announced (its stage frees later); expect_tx may precede the copy.
      const int nk0 = n_tiles < 2 ? n_tiles : 2;
      mbar_expect_tx(q_full, Cfg::Q_BYTES + nk0 * Cfg::KV_BYTES);
#pragma unroll
      for (int s = 0; s < KS; ++s)
        tma_load_4d(sm + (sQ - base) + s * (ATC_BQ * 128), &p.mapQ, q_full, s * 64, h, qt * ATC_BQ, b);
      for (int j = 0; j < nk0; ++j) load_k(j, q_full);
      for (int t = 0; t < 2 && t < n_tiles; ++t) {
        mbar_expect_tx(rdy + t, Cfg::KV_BYTES + (t + 2 < n_tiles ? Cfg::KV_BYTES : 0));
        load_v(t, rdy + t);
        if (t + 2 < n_tiles && t + 2 < NST) load_k(t + 2, rdy + t);
      }
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(rdy + (j & 1), (uint32_t)((j >> 1) & 1));  // P_j written, S_j read, V_j and K_{j+2} landed
        tc_fence_after();
        issue_pv(j);
        if (j == n_tiles - 1) umma_commit(fin);
        if (j + 2 < n_tiles) {
          issue_s(j + 2);
          // this thread's arrival for tile j + 2, with the bytes its wake-up will need
          mbar_expect_tx(rdy + (j & 1), Cfg::KV_BYTES + (j + 4 < n_tiles ? Cfg::KV_BYTES : 0));
          // V_{j+2} takes the stage of V_{j+2-NST}: P V_{j-2} finished before S_j was issued (NST = 4); with NST = 3 it
          // is P V_{j-1}, issued one wake-up ago: wait for it (cheap: it ran during the whole softmax of tile j)
          if (NST == 3 && j >= 1) mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
          load_v(j + 2, rdy + (j & 1));
        }
        if (NST >= 4) {
          if (j + 4 < n_tiles) load_k(j + 4, rdy + (j & 1));        // K_j's stage: S_j has been read
        } else {
          if (j + 3 < n_tiles) load_k(j + 3, rdy + ((j + 1) & 1));  // K_j's stage; bytes announced at the last wake-up
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax warps: thread = query row
    const int r = tid;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl = p.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    const int rx = r & 7;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tS = tmem + (uint32_t)(j & 1) * ATC_BKV;
      mbar_wait(s_full + (j & 1), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t sv[2][32];
      tmem_ld_32x32(tS + lane_off, sv[0]);
      tmem_ld_32x32(tS + lane_off + 32, sv[1]);
      tmem_ld_wait();
      float* s = reinterpret_cast<float*>(&sv[0][0]);
      const int kbase = j * ATC_BKV;
      if constexpr (PACK) {
        // block-diagonal mask: row r = pixel (r >> flog) attends to the keys of the same pixel only
        const int my_blk = r >> p.pack_flog;
#pragma unroll
        for (int c = 0; c < ATC_BKV; ++c)
          if (((kbase + c) >> p.pack_flog) != my_blk) s[c] = ATC_PACK_NEG;
      } else if (kbase + ATC_BKV > p.skv) {
#pragma unroll
        for (int c = 0; c < ATC_BKV; ++c)
          if (kbase + c >= p.skv) s[c] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < ATC_BKV; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const bool grow = (mx - m_ref) * sl > ATC_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_ref, mx);
        const float alpha = ex2_approx((m_ref - m_new) * sl);
        m_ref = m_new;
        l *= alpha;
        if (j > 0) {
          mbar_wait(pv_done, (uint32_t)((j - 1) & 1));  // O holds tiles 0..j-1 (rare path)
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < DP; c += 16) {
            uint32_t ov[16];
            tmem_ld_32x16(tO + lane_off + c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(tO + lane_off + c, ov);
          }
          tmem_st_wait();
        }
      }
      const float mb = m_ref * sl;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[ATC_BKV / 2];
#pragma unroll
      for (int c = 0; c < ATC_BKV; c += 2) {
        const float e0 = ex2_approx(fmaf(s[c], sl, -mb));
        const float e1 = ex2_approx(fmaf(s[c + 1], sl, -mb));
        sum0 += e0;
        sum1 += e1;
        pk[c >> 1] = pack_half2(e0, e1);
      }
      l += sum0 + sum1;
      // P buffer j & 1 was read by P V_{j-2}, which completed before S_j was issued: free
      const uint32_t p_row = sP + (uint32_t)(j & 1) * Cfg::P_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < ATC_BKV / 8; ++c)
        st_shared_v4(p_row + ((c ^ rx) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(rdy + (j & 1));
    }
    // own barrier for the end: a warp that only knows "P V_{n-3} is complete" cannot name the phase of pv_done it needs
    // by parity (n-1 aliases with n-3 while n-2 runs; n-2 aliases with the never-completing phase n once both are done)
    mbar_wait(fin, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    int qrow = qt * ATC_BQ + r;
    __half* dst;
    if constexpr (PACK) {
      // row r = (pixel r >> flog, frame r & (F - 1)); o_batch_stride = rows per batch element, o_seq_stride = rows per
      // frame, pixels are consecutive rows
      const int pix = qt * ppt + (r >> p.pack_flog);
      const int fr = r & ((1 << p.pack_flog) - 1);
      dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)fr * p.o_seq_stride + pix) * p.ldo + h * D;
      qrow = (pix < p.pack_hw) ? 0 : p.sq;  // the guard below: rows of pixels past the image are not stored
    } else {
      dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)qrow * p.o_seq_stride) * p.ldo + h * D;
    }
#pragma unroll
    for (int c = 0; c < DP; c += 16) {
      uint32_t ov[16];
      tmem_ld_32x16(tO + lane_off + c, ov);
      tmem_ld_wait();
      if (qrow < p.sq) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < D) {
            uint4 pk4;
            pk4.x = pack_half2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
            pk4.y = pack_half2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            pk4.z = pack_half2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
            pk4.w = pack_half2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + i) = pk4;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
// =================================================================================================================
// One-tile attention: every problem whose keys fit ONE key tile - the audio cross-attention (50 keys), the 8x8 / 4x4
// levels (64 / 16 keys) and, PACK = true, the temporal attention over F <= 32 frames (a 128-row tile of 128 / F pixels x
// F frames attends to itself through a block-diagonal mask).  No online softmax, no rescale: S = Q K^T, one softmax pass,
// O = P V.  The flash kernel above pays ~8-10 k clocks of per-CTA setup and hand-shake latency, fine for 16 key tiles and
// ruinous for one (packed temporal attention: 40 us against 24 us on the warp-level kernel), so this kernel is
// PERSISTENT: a CTA walks over (query tile, head) work items, its control thread keeps the Q / K / V loads of the next
// items in flight (ring of NST stages) and issues S_{i+1} while the four softmax warps (thread = row) are still busy
// with item i; O is double buffered in TMEM so that the store of item i - 1 overlaps P V_i.
//   per item:  control  wait loads_i, [S_{i-1} read] -> S_i MMAs -> (S_i done => P V_{i-1} done: refill its stage)
//                       wait P_i, [O buffer read]   -> P V_i MMAs
//              warps    wait S_i -> tcgen05.ld -> mask / max / exp2 / sum -> P_i (fp16, 128B-swizzled A tile) -> signal
//                       then the epilogue of item i - 1: tcgen05.ld O, * 1 / l, fp16 rows to global
// PACK: the row's keys are columns [32 w, 32 w + 32) of S for warp w (rows are pixel-major, F divides 32), so a thread
// loads ONE 32-column block, masks the other pixels' columns, and writes 64 bytes of P; the rest of the P tile is zeroed
// once per CTA and never touched again.
// =================================================================================================================
template <int D, bool PACK>
struct AtoCfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int KS = (DP + 63) / 64;
  static constexpr int KSTEPS = DP / 16;
  static constexpr int KT = PACK ? 128 : 64;             // keys per tile
  static constexpr int Q_BYTES = KS * ATC_BQ * 128;
  static constexpr int KV_SLAB = KT * 128;
  static constexpr int KV_BYTES = KS * KV_SLAB;
  static constexpr int STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;
  static constexpr int P_BYTES = (KT / 64) * ATC_BQ * 128;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int NST_RAW = (BUDGET - P_BYTES) / STAGE_BYTES;
  static constexpr int NST = NST_RAW > 3 ? 3 : (NST_RAW < 1 ? 1 : NST_RAW);
  static constexpr int TILE_BYTES = NST * STAGE_BYTES + P_BYTES;
  static constexpr int SMEM = TILE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
  static constexpr int TMEM_NEED = KT + 2 * DP;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int CTAS_PER_SM = (2 * SMEM <= 226 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
};

struct AttnOneParams {
  CUtensorMap mapQ, mapK, mapV;
  __half* o;
  int ldo;
  int sq, skv;
  int64_t o_batch_stride;  // rows
  int64_t o_seq_stride;    // rows
  float scale_log2;
  int heads, tiles_q;      // work item w -> h = w % heads, qt = (w / heads) % tiles_q, b = w / (heads * tiles_q)
  int n_items;
  int pack_flog, pack_hw;
};

template <int D, bool PACK>
__global__ void __launch_bounds__(ATC_THREADS, AtoCfg<D, PACK>::CTAS_PER_SM) attn_one_kernel(const __grid_constant__ AttnOneParams p) {
  using Cfg = AtoCfg<D, PACK>;
  constexpr int DP = Cfg::DP, KS = Cfg::KS, KSTEPS = Cfg::KSTEPS, KT = Cfg::KT, NST = Cfg::NST;
  pdl_prologue();
  extern __shared__ uint8_t ato_smem_raw[];
  const uint32_t raw = smem_u32(ato_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = ato_smem_raw + (base - raw);
  const uint32_t sP = base + NST * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::TILE_BYTES);
  uint64_t* full = bars + 0;      // [NST] loads of an item have landed
  uint64_t* s_full = bars + 4;    // S_i complete (and every earlier MMA)
  uint64_t* s_free = bars + 5;    // the 4 warps hold S_i in registers
  uint64_t* p_full = bars + 6;    // the 4 warps have written P_i
  uint64_t* o_full = bars + 7;    // [2] P V_i complete
  uint64_t* o_free = bars + 9;    // [2] the 4 warps have read O buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mine = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items w = blockIdx.x + i * gridDim.x
  const int ppt = PACK ? (ATC_BQ >> p.pack_flog) : 0;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(full + s, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full + 0, 1);
    mbar_init(o_full + 1, 1);
    mbar_init(o_free + 0, 4);
    mbar_init(o_free + 1, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem;         // columns [0, KT): scores
  const uint32_t tO = tmem + KT;    // columns [KT, KT + 2 DP): two output accumulators

  auto decode = [&](int i, int& qt, int& h, int& b) {
    const int w = (int)blockIdx.x + i * (int)gridDim.x;
    h = w % p.heads;
    const int t = w / p.heads;
    qt = t % p.tiles_q;
    b = t / p.tiles_q;
  };

  if (warp == 4) {
    if (elect_one()) {
      // ------------------------------------------------------------------ control: TMA producer + MMA issuer
      tma_prefetch_desc(&p.mapQ);
      tma_prefetch_desc(&p.mapK);
      tma_prefetch_desc(&p.mapV);
      constexpr uint32_t idesc_s = (1u << 4) | (uint32_t(KT >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 16) | (uint32_t(DP >> 3) << 17) | (uint32_t(ATC_BQ >> 4) << 24);
      auto load_item = [&](int i) {
        int qt, h, b;
        decode(i, qt, h, b);
        const int st = i % NST;
        uint8_t* q_dst = sm + st * Cfg::STAGE_BYTES;
        uint8_t* k_dst = q_dst + Cfg::Q_BYTES;
        uint8_t* v_dst = k_dst + Cfg::KV_BYTES;
        mbar_expect_tx(full + st, Cfg::STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if constexpr (PACK) {
            tma_load_5d(q_dst + s * (ATC_BQ * 128), &p.mapQ, full + st, s * 64, h, 0, qt * ppt, b);
            tma_load_5d(k_dst + s * Cfg::KV_SLAB, &p.mapK, full + st, s * 64, h, 0, qt * ppt, b);
            tma_load_5d(v_dst + s * Cfg::KV_SLAB, &p.mapV, full + st, s * 64, h, 0, qt * ppt, b);
          } else {
            tma_load_4d(q_dst + s * (ATC_BQ * 128), &p.mapQ, full + st, s * 64, h, qt * ATC_BQ, b);
            tma_load_4d(k_dst + s * Cfg::KV_SLAB, &p.mapK, full + st, s * 64, h, 0, b);
            tma_load_4d(v_dst + s * Cfg::KV_SLAB, &p.mapV, full + st, s * 64, h, 0, b);
          }
        }
      };
      for (int i = 0; i < NST && i < n_mine; ++i) load_item(i);
      for (int i = 0; i < n_mine; ++i) {
        const int st = i % NST;
        const uint32_t sQ = base + st * Cfg::STAGE_BYTES;
        const uint32_t sK = sQ + Cfg::Q_BYTES;
        const uint32_t sV = sK + Cfg::KV_BYTES;
        mbar_wait(full + st, (uint32_t)((i / NST) & 1));
        if (i > 0) mbar_wait(s_free, (uint32_t)((i - 1) & 1));  // the warps hold S_{i-1}: its TMEM columns are free
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t qa = sQ + (ks >> 2) * (ATC_BQ * 128) + (ks & 3) * 32;
          const uint32_t ka = sK + (ks >> 2) * Cfg::KV_SLAB + (ks & 3) * 32;
          umma_f16_ss(tS, umma_desc_sw128(qa), umma_desc_sw128(ka), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        if (NST > 1) {
          // S_i complete => P V_{i-1} complete (the tensor pipe runs in issue order): stage (i - 1) % NST can be refilled
          mbar_wait(s_full, (uint32_t)(i & 1));
          if (i >= 1 && i - 1 + NST < n_mine) load_item(i - 1 + NST);
        }
        mbar_wait(p_full, (uint32_t)(i & 1));
        if (i >= 2) mbar_wait(o_free + (i & 1), (uint32_t)(((i - 2) >> 1) & 1));  // the epilogue of item i - 2 has read this O
        tc_fence_after();
        const uint32_t tOi = tO + (uint32_t)(i & 1) * DP;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint32_t pa = sP + (kk >> 2) * (ATC_BQ * 128) + (kk & 3) * 32;
          const uint32_t va = sV + kk * 16 * 128;
          umma_f16_ss(tOi, umma_desc_sw128(pa), umma_desc_mn_sw128(va, Cfg::KV_SLAB), idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_full + (i & 1));
        if (NST == 1 && i + 1 < n_mine) {
          mbar_wait(o_full + (i & 1), (uint32_t)((i >> 1) & 1));  // single stage: V_i is read by P V_i
          load_item(i + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------------- softmax + epilogue: thread = row
    const int r = tid;  // 0..127 = TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sl = p.scale_log2;
    const int rx = r & 7;
    if constexpr (PACK) {
      // zero this row of the P tile once: only the row's own 32-column block is ever written afterwards
#pragma unroll
      for (int a = 0; a < KT / 64; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) st_shared_v4(sP + a * (ATC_BQ * 128) + r * 128 + (c << 4), 0u, 0u, 0u, 0u);
    }
    float inv_prev = 0.f;
    auto epilogue = [&](int i, float inv) {
      int qt, h, b;
      decode(i, qt, h, b);
      mbar_wait(o_full + (i & 1), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t tOi = tO + (uint32_t)(i & 1) * DP + lane_off;
      int qrow = qt * ATC_BQ + r;
      __half* dst;
      if constexpr (PACK) {
        const int pix = qt * ppt + (r >> p.pack_flog);
        const int fr = r & ((1 << p.pack_flog) - 1);
        dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)fr * p.o_seq_stride + pix) * p.ldo + h * D;
        qrow = (pix < p.pack_hw) ? 0 : p.sq;
      } else {
        dst = p.o + ((int64_t)b * p.o_batch_stride + (int64_t)qrow * p.o_seq_stride) * p.ldo + h * D;
      }
#pragma unroll
      for (int c = 0; c < DP; c += 16) {
        uint32_t ov[16];
        tmem_ld_32x16(tOi + c, ov);
        tmem_ld_wait();
        if (c + 16 >= DP) {  // last chunk is in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free + (i & 1));
        }
        if (qrow < p.sq) {
#pragma unroll
          for (int k = 0; k < 16; k += 8) {
            if (c + k < D) {
              uint4 pk4;
              pk4.x = pack_half2(__uint_as_float(ov[k]) * inv, __uint_as_float(ov[k + 1]) * inv);
              pk4.y = pack_half2(__uint_as_float(ov[k + 2]) * inv, __uint_as_float(ov[k + 3]) * inv);
              pk4.z = pack_half2(__uint_as_float(ov[k + 4]) * inv, __uint_as_float(ov[k + 5]) * inv);
              pk4.w = pack_half2(__uint_as_float(ov[k + 6]) * inv, __uint_as_float(ov[k + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + c + k) = pk4;
            }
          }
        }
      }
    };
    for (int i = 0; i < n_mine; ++i) {
      mbar_wait(s_full, (uint32_t)(i & 1));
      tc_fence_after();
      float l = 0.f;
      if constexpr (PACK) {
        uint32_t sv[32];
        tmem_ld_32x32(tS + lane_off + 32 * warp, sv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float* s = reinterpret_cast<float*>(sv);
        const int my_blk = r >> p.pack_flog;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (((32 * warp + c) >> p.pack_flog) != my_blk) s[c] = -INFINITY;
        float mx = s[0];
#pragma unroll
        for (int c = 1; c < 32; ++c) mx = fmaxf(mx, s[c]);
        const float mb = mx * sl;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float e0 = ex2_approx(fmaf(s[c], sl, -mb));
          const float e1 = ex2_approx(fmaf(s[c + 1], sl, -mb));
          l += e0 + e1;
          pk[c >> 1] = pack_half2(e0, e1);
        }
        // 32 keys = 64 bytes = chunks (warp & 1) * 4 .. + 3 of the row in 64-key atom (warp >> 1)
        const uint32_t p_row = sP + (warp >> 1) * (ATC_BQ * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_shared_v4(p_row + ((((warp & 1) * 4 + c) ^ rx) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      } else {
        uint32_t sv[2][32];
        tmem_ld_32x32(tS + lane_off, sv[0]);
        tmem_ld_32x32(tS + lane_off + 32, sv[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float* s = reinterpret_cast<float*>(&sv[0][0]);
        if (p.skv < KT) {
#pragma unroll
          for (int c = 0; c < KT; ++c)
            if (c >= p.skv) s[c] = -INFINITY;
        }
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int c = 4; c < KT; c += 4) {
          mx0 = fmaxf(mx0, s[c]);
          mx1 = fmaxf(mx1, s[c + 1]);
          mx2 = fmaxf(mx2, s[c + 2]);
          mx3 = fmaxf(mx3, s[c + 3]);
        }
        const float mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl;
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pk[KT / 2];
#pragma unroll
        for (int c = 0; c < KT; c += 2) {
          const float e0 = ex2_approx(fmaf(s[c], sl, -mb));
          const float e1 = ex2_approx(fmaf(s[c + 1], sl, -mb));
          sum0 += e0;
          sum1 += e1;
          pk[c >> 1] = pack_half2(e0, e1);
        }
        l = sum0 + sum1;
        const uint32_t p_row = sP + r * 128;
#pragma unroll
        for (int c = 0; c < KT / 8; ++c)
          st_shared_v4(p_row + ((c ^ rx) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (i > 0) epilogue(i - 1, inv_prev);
      inv_prev = 1.f / l;
    }
    if (n_mine > 0) epilogue(n_mine - 1, inv_prev);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}


typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled atc_encode_fn() {
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

// [batch][seq][head][d] view of a token matrix: element (b, s, h, c) at ptr + ((b*outer + s*seq) * ld + h*D + c)
static int encode_bshd(PFN_tmapEncodeTiled encode, CUtensorMap* map, const void* ptr, int D, int heads, int seq,
                       int batch, int64_t ld, int64_t seq_stride, int64_t outer_stride, int box_rows, const char* what) {
  cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)heads, (cuuint64_t)seq, (cuuint64_t)batch};
  cuuint64_t gstr[3] = {(cuuint64_t)D * 2, (cuuint64_t)(seq_stride * ld) * 2, (cuuint64_t)(outer_stride * ld) * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LS_CHECK(r == CUDA_SUCCESS, "ls_attention: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return 0;
}

static bool atc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LS_ATTN_TC");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

#ifdef LS_ATC_PROBE
static long long* g_atc_probe = nullptr;
extern "C" int ls_atc_probe_read(long long* host, int n) {
  if (!g_atc_probe) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(host, g_atc_probe, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
#endif

static int atc_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LS_ATTN_V");
    v = e ? atoi(e) : 2;
  }
  return v;
}

// PACK maps: [d][head][frame][pixel][batch element] view of the token matrix [(b f) (h w)] x ld
static int encode_pack(PFN_tmapEncodeTiled encode, CUtensorMap* map, const void* ptr, int D, int heads, int F, int hw,
                       int nb, int64_t ld, int64_t seq_stride, int64_t inner_stride, int64_t outer_stride, int box_pix,
                       const char* what) {
  cuuint64_t gdim[5] = {(cuuint64_t)D, (cuuint64_t)heads, (cuuint64_t)F, (cuuint64_t)hw, (cuuint64_t)nb};
  cuuint64_t gstr[4] = {(cuuint64_t)D * 2, (cuuint64_t)(seq_stride * ld) * 2, (cuuint64_t)(inner_stride * ld) * 2,
                        (cuuint64_t)(outer_stride * ld) * 2};
  cuuint32_t box[5] = {64, 1, (cuuint32_t)F, (cuuint32_t)box_pix, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LS_CHECK(r == CUDA_SUCCESS, "ls_attention: cuTensorMapEncodeTiled(%s, packed) failed with %d", what, (int)r);
  return 0;
}

template <int D, bool PACK>
static int launch_attn_tc(const LsAttnArgs* a, cudaStream_t stream) {
  using Cfg = AtcCfg<D>;
  PFN_tmapEncodeTiled encode = atc_encode_fn();
  LS_CHECK(encode != nullptr, "ls_attention: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  dim3 grid;
  if constexpr (PACK) {
    // temporal attention: batch = nb * hw sequences of F tokens; sequence (e, pix) lives at rows e * outer + pix * inner
    // (+ f * seq).  Tile = 128 / F pixels x F frames.
    const int F = a->sq, hw = a->q_inner, nb = a->batch / a->q_inner;
    int flog = 0;
    while ((1 << flog) < F) ++flog;
    const int ppt = ATC_BQ / F, ppk = ATC_BKV / F;
    if (encode_pack(encode, &p.mapQ, a->q, D, a->heads, F, hw, nb, a->ldq, a->q_seq_stride, a->q_inner_stride,
                    a->q_outer_stride, ppt, "Q") ||
        encode_pack(encode, &p.mapK, a->k, D, a->heads, F, hw, nb, a->ldk, a->kv_seq_stride, a->kv_inner_stride,
                    a->kv_outer_stride, ppk, "K") ||
        encode_pack(encode, &p.mapV, a->v, D, a->heads, F, hw, nb, a->ldv, a->kv_seq_stride, a->kv_inner_stride,
                    a->kv_outer_stride, ppk, "V"))
      return 1;
    p.pack_flog = flog;
    p.pack_hw = hw;
    grid = dim3((hw + ppt - 1) / ppt, a->heads, nb);
  } else {
    if (encode_bshd(encode, &p.mapQ, a->q, D, a->heads, a->sq, a->batch, a->ldq, a->q_seq_stride, a->q_outer_stride,
                    ATC_BQ, "Q"))
      return 1;
    if (encode_bshd(encode, &p.mapK, a->k, D, a->heads, a->skv, a->batch, a->ldk, a->kv_seq_stride,
                    a->kv_outer_stride, ATC_BKV, "K"))
      return 1;
    if (encode_bshd(encode, &p.mapV, a->v, D, a->heads, a->skv, a->batch, a->ldv, a->kv_seq_stride,
                    a->kv_outer_stride, ATC_BKV, "V"))
      return 1;
    grid = dim3((a->sq + ATC_BQ - 1) / ATC_BQ, a->heads, a->batch);
  }
  p.o = reinterpret_cast<__half*>(a->out);
  p.ldo = a->ldo;
  p.sq = a->sq;
  p.skv = a->skv;
  p.o_batch_stride = a->q_outer_stride;
  p.o_seq_stride = a->q_seq_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
#ifdef LS_ATC_PROBE
  {
    static long long* probe_buf = nullptr;
    if (!probe_buf) {
      cudaMalloc(&probe_buf, 16 * 64 * 16 * sizeof(long long));
      cudaMemset(probe_buf, 0, 16 * 64 * 16 * sizeof(long long));
    }
    p.probe = probe_buf;
    g_atc_probe = probe_buf;
  }
#endif
  static bool attr_set_dev[16] = {};
  bool& attr_set = attr_set_dev[dev_slot()];
  if (!attr_set) {
    LS_CUDA(cudaFuncSetAttribute(attn_tc_kernel<D, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    if (D == 80 && !PACK)
      LS_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Atc2Cfg<D>::SMEM));
    attr_set = true;
  }
  // Measured (tools/attn_bench.py): version 2 wins at head_dim 80 (S = 1024: 540 -> 419 us, S = 256: 33.5 -> 31.0 us)
  // and ties at head_dim 40 (129.6 vs 127.8 us: with one or four waits per tile, two or three CTAs per SM, both land at
  // ~1100 clocks per key tile per SM - the softmax warps' own per-tile latency chain, not the control thread, is
  // what is left.  A third variant with 8 softmax warps, two threads per query row, also ties: 130.6 us at head_dim
  // 40, 432 us at head_dim 80 / S = 1024 (tools/experiments/r1_attention_v3.patch) - so neither the control thread nor
  // the warps' latency chain is the limit; the candidates left are shared-memory bandwidth (per key tile the MMAs
  // re-read 16 KB of Q and 16 KB of P, the warps write 16 KB of P, TMA writes 16 KB of K/V: ~80 KB against 128 B/clk)
  // and the TMEM read of S (32 KB per tile).  Next: P as the A operand from TMEM.)  head_dim 160 stays on version 1 (the
  // 3-stage ring variant of version 2 still has an accounting bug: its control thread times out).  Version 2 only takes
  // the long self-attention sequences it was written for (>= 2 full key tiles).
  if (!PACK && atc_version() == 2 && D == 80 && a->sq >= ATC_BQ && a->skv >= 2 * ATC_BKV) {
    static const int pad = getenv("LS_ATTN_SMEM_PAD") ? atoi(getenv("LS_ATTN_SMEM_PAD")) : 0;  // debugging: forces 1 CTA/SM
    size_t smem2 = (size_t)Atc2Cfg<D>::SMEM + (size_t)pad;
    if (smem2 > 227 * 1024) smem2 = 227 * 1024;
    if (pad) cudaFuncSetAttribute(attn_tc2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    LS_CUDA(launch_k(attn_tc2_kernel<D>, grid, dim3(ATC_THREADS), smem2, stream, p));
  }
  else {
    LS_CUDA(launch_k(attn_tc_kernel<D, PACK>, grid, dim3(ATC_THREADS), (size_t)Cfg::SMEM, stream, p));
  }
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <int D>
static int launch_attn_tc_any(const LsAttnArgs* a, bool pack, cudaStream_t stream) {
  return pack ? launch_attn_tc<D, true>(a, stream) : launch_attn_tc<D, false>(a, stream);
}

template <int D, bool PACK>
static int launch_attn_one(const LsAttnArgs* a, cudaStream_t stream) {
  using Cfg = AtoCfg<D, PACK>;
  PFN_tmapEncodeTiled encode = atc_encode_fn();
  LS_CHECK(encode != nullptr, "ls_attention: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  AttnOneParams p;
  memset(&p, 0, sizeof(p));
  int nb;
  if constexpr (PACK) {
    const int F = a->sq, hw = a->q_inner;
    nb = a->batch / a->q_inner;
    int flog = 0;
    while ((1 << flog) < F) ++flog;
    const int ppt = ATC_BQ / F;
    if (encode_pack(encode, &p.mapQ, a->q, D, a->heads, F, hw, nb, a->ldq, a->q_seq_stride, a->q_inner_stride,
                    a->q_outer_stride, ppt, "Q") ||
        encode_pack(encode, &p.mapK, a->k, D, a->heads, F, hw, nb, a->ldk, a->kv_seq_stride, a->kv_inner_stride,
                    a->kv_outer_stride, ppt, "K") ||
        encode_pack(encode, &p.mapV, a->v, D, a->heads, F, hw, nb, a->ldv, a->kv_seq_stride, a->kv_inner_stride,
                    a->kv_outer_stride, ppt, "V"))
      return 1;
    p.pack_flog = flog;
    p.pack_hw = hw;
    p.tiles_q = (hw + ppt - 1) / ppt;
  } else {
    nb = a->batch;
    if (encode_bshd(encode, &p.mapQ, a->q, D, a->heads, a->sq, a->batch, a->ldq, a->q_seq_stride, a->q_outer_stride,
                    ATC_BQ, "Q") ||
        encode_bshd(encode, &p.mapK, a->k, D, a->heads, a->skv, a->batch, a->ldk, a->kv_seq_stride,
                    a->kv_outer_stride, Cfg::KT, "K") ||
        encode_bshd(encode, &p.mapV, a->v, D, a->heads, a->skv, a->batch, a->ldv, a->kv_seq_stride,
                    a->kv_outer_stride, Cfg::KT, "V"))
      return 1;
    p.tiles_q = (a->sq + ATC_BQ - 1) / ATC_BQ;
  }
  p.o = reinterpret_cast<__half*>(a->out);
  p.ldo = a->ldo;
  p.sq = a->sq;
  p.skv = a->skv;
  p.o_batch_stride = a->q_outer_stride;
  p.o_seq_stride = a->q_seq_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.heads = a->heads;
  const int64_t n_items = (int64_t)nb * p.tiles_q * a->heads;
  LS_CHECK(n_items < (1ll << 30), "ls_attention: too many work items");
  p.n_items = (int)n_items;
  static bool attr_set_dev[16] = {};
  bool& attr_set = attr_set_dev[dev_slot()];
  if (!attr_set) {
    LS_CUDA(cudaFuncSetAttribute(attn_one_kernel<D, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  int sms = 0;
  LS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_slot()));
  const int64_t slots = (int64_t)(sms > 0 ? sms : 148) * Cfg::CTAS_PER_SM;
  const int grid = (int)(n_items < slots ? n_items : slots);
  LS_CUDA(launch_k(attn_one_kernel<D, PACK>, dim3(grid), dim3(ATC_THREADS), (size_t)Cfg::SMEM, stream, p));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <int D>
static int launch_attn_one_any(const LsAttnArgs* a, bool pack, cudaStream_t stream) {
  return pack ? launch_attn_one<D, true>(a, stream) : launch_attn_one<D, false>(a, stream);
}

// returns -1 when the problem is not one for this path (the caller then uses the warp-level kernels)
int attention_tc_try(const LsAttnArgs* a, cudaStream_t stream) {
  if (!atc_enabled()) return -1;
  // Routing.  Long key sequences (>= 2 key tiles: the 32x32 / 16x16 spatial self-attention) take the flash kernel.  The
  // problems whose keys fit one tile - audio cross-attention, 8x8 / 4x4 levels, temporal attention - have two tcgen05
  // implementations here (LS_ATTN_ONE=1: the persistent one-tile kernel; LS_ATTN_TC_ALL=1: the flash kernel, packed
  // temporal mode), both parity-green and both SLOWER than the warp-level kernels of attention.cu, which stay the
  // default: level-0 temporal attention 24.7 us (mma.sync) vs 36-41 us (either tcgen05 kernel), 4x4 level 8.5 vs 12.9 us,
  // audio cross-attention 20.8 vs 21.0 us (profiles/r2d_attention_all_tcgen05.txt, r2g_attention_one_tile.txt).  ncu on
  // the one-tile kernel (profiles/r2g_attn_one_ncu_summary.txt): tensor pipe 10 % active, the softmax warps wait for S a
  // third of the time, 1 600 L2 requests per (8 pixels x 1 head) item - these problems are bound by gathering 80-byte
  // rows (q, k, v of one head out of a 1 920-byte token row), which a TMA box does no better than cp.async, and their
  // 12 GFLOP are noise for any tensor path.
  static int all = -1, one = -1;
  if (all < 0) {
    const char* e = getenv("LS_ATTN_TC_ALL");
    all = (e && e[0] == '1') ? 1 : 0;
    const char* o = getenv("LS_ATTN_ONE");
    one = (o && o[0] == '1') ? 1 : 0;
  }
  bool pack = false;
  if (a->q_inner != 1 || a->kv_inner != 1) {
    // strided batches = temporal attention: packed mode needs self-attention geometry (same addressing for q and k/v),
    // consecutive pixels one row apart, and a power-of-two sequence length that divides a 64-key tile
    const int F = a->sq;
    if (a->sq != a->skv || F < 2 || F > ATC_BKV || (F & (F - 1)) != 0) return -1;
    if (a->q_inner != a->kv_inner || a->q_inner_stride != 1 || a->kv_inner_stride != 1) return -1;
    if (a->q_outer_stride != a->kv_outer_stride || a->q_seq_stride != a->kv_seq_stride) return -1;
    if (a->batch % a->q_inner != 0) return -1;
    if (a->batch / a->q_inner > 65535) return -1;
    pack = true;
  } else {
    if (a->batch > 65535) return -1;
  }
  if (a->heads > 65535) return -1;
  if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15)
    return -1;
  if (a->head_dim % 8) return -1;  // 16-byte rows per head for TMA and the vector stores
  const bool long_keys = !pack && a->sq >= ATC_BQ && a->skv >= 2 * ATC_BKV;
  const bool one_tile = pack ? (a->sq <= 32) : (a->skv <= ATC_BKV);
  if (!long_keys && !all) {
    if (!one || !one_tile) return -1;
    switch (a->head_dim) {
      case 16: return launch_attn_one_any<16>(a, pack, stream);
      case 32: return launch_attn_one_any<32>(a, pack, stream);
      case 40: return launch_attn_one_any<40>(a, pack, stream);
      case 64: return launch_attn_one_any<64>(a, pack, stream);
      case 80: return launch_attn_one_any<80>(a, pack, stream);
      case 160: return launch_attn_one_any<160>(a, pack, stream);
      default: return -1;
    }
  }
  switch (a->head_dim) {
    case 16: return launch_attn_tc_any<16>(a, pack, stream);
    case 32: return launch_attn_tc_any<32>(a, pack, stream);
    case 40: return launch_attn_tc_any<40>(a, pack, stream);
    case 64: return launch_attn_tc_any<64>(a, pack, stream);
    case 80: return launch_attn_tc_any<80>(a, pack, stream);
    case 160: return launch_attn_tc_any<160>(a, pack, stream);
    default: return -1;
  }
}

}  // namespace ls
