// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (fp16 x fp16 -> fp32 in TMEM)
//   warps 2..9  : epilogue       (tcgen05.ld -> bias / time-embedding / residual / GEGLU / SiLU -> 64B-swizzled smem
//                                 slab -> TMA bulk store; two groups of 4 warps take alternate 32-column slabs)
//
// Two TMEM accumulators (double buffered) let the epilogue of tile i overlap the main loop of tile i+1.
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

namespace ls {

extern std::atomic<int64_t> g_launch_count;

constexpr int BM = 128;
constexpr int BK = 64;  // fp16 elements: one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 320;
constexpr int SLAB_BYTES = BM * 64;      // 128 rows x 32 fp16 columns, SWIZZLE_64B
constexpr int STAGING_BYTES = 4 * SLAB_BYTES;  // 2 epilogue groups x 2 slabs (ping-pong)
constexpr int MAX_STAGES = 8;

struct GemmKParams {
  CUtensorMap mapA[LS_GEMM_MAX_SEG];
  CUtensorMap mapB;
  int nseg;
  int seg_taps[LS_GEMM_MAX_SEG];
  int seg_cblk[LS_GEMM_MAX_SEG];
  int bw, bh, bn;  // TMA box in pixels: bw * bh * bn == 128
  int H, W, nimg;
  int tiles_x, tiles_y;
  int m_tiles, n_tiles, num_kb;
  int N;
  int b_batched;
  int stages;
  const float* bias;
  int bias_div;
  int bias_ld;
  const __half* residual;
  int ldr;
  void* out;
  int ldo;
  int flags;
  CUtensorMap mapOut;  // fp16 [M][N_out] output, box 32 columns x 128 rows, SWIZZLE_64B (valid iff tma_store)
  int tma_store;
  int64_t M;           // total output rows
};

__device__ __forceinline__ void decode_m_tile(const GemmKParams& p, int mt, int& x0, int& y0, int& i0) {
  const int tx = mt % p.tiles_x;
  const int rest = mt / p.tiles_x;
  const int ty = rest % p.tiles_y;
  const int tn = rest / p.tiles_y;
  x0 = tx * p.bw;
  y0 = ty * p.bh;
  i0 = tn * p.bn;
}

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
      for (int j = 0; j < nvalid; ++j) f[j] += __half2float(rr[j]);
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
      for (int j = 0; j < nvalid; ++j) o[j] = f[j];
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
      for (int j = 0; j < nvalid; ++j) o[j] = __float2half_rn(f[j]);
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
    for (int j = 0; j < nvalid; ++j) f[j] += __ldg(b + j);
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "tile N");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment for SWIZZLE_128B");
  // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24
  constexpr uint32_t IDESC = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int stages = p.stages;
  uint8_t* staging = smem + stages * STAGE_BYTES;  // 1024-byte aligned (STAGE_BYTES % 1024 == 0)
  float* bias_sm = reinterpret_cast<float*>(staging + STAGING_BYTES);  // 2 groups x 256 floats
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES + 2048);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.mapA[s]);
    tma_prefetch_desc(&p.mapB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], GEMM_THREADS / 32 - 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int nt = tile - mt * p.n_tiles;
        int x0, y0, i0;
        decode_m_tile(p, mt, x0, y0, i0);
        const int bz = p.b_batched ? i0 : 0;
        int kcol = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          const int cblk = p.seg_cblk[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            for (int cb = 0; cb < cblk; ++cb) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              uint8_t* sa = smem + stage * STAGE_BYTES;
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              tma_load_4d(sa, &p.mapA[s], &full_bar[stage], cb * BK, x0 + dx, y0 + dy, i0);
              tma_load_3d(sa + A_STAGE_BYTES, &p.mapB, &full_bar[stage], kcol, nt * BN, bz);
              kcol += BK;
              if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const uint32_t acc_phase = (lt >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes along K inside the 128B swizzle atom == +2 in the 16-byte-unit address field
            umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_kb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int group = (warp - 2) >> 2; // 0: even slabs, 1: odd slabs
    const int r = q * 32 + lane;       // row of the tile == TMEM lane
    const bool geglu = (p.flags & LS_EPI_GEGLU) != 0;
    if (p.tma_store) {
      // Latency-tolerant epilogue: the tile's bias row goes to smem and ALL residual fragments of the group's slabs
      // are requested before waiting for the accumulator, so global-load latency hides behind the main loop.
      const bool issuer = (warp == 2 + 4 * group) && lane == 0;
      uint8_t* my_stage = staging + group * 2 * SLAB_BYTES;
      float* my_bias = bias_sm + group * 256;
      constexpr int NSLAB_MAX = (BN / 32 + 1) / 2;  // slabs per group
      const int nslab = geglu ? BN / 64 : BN / 32;
      const int n_out_total = geglu ? p.N / 2 : p.N;
      const int gtid = (warp - 2 - 4 * group) * 32 + lane;  // 0..127 inside the group
      uint32_t slab_count = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
        const int mt = tile / p.n_tiles;
        const int nt = tile - mt * p.n_tiles;
        int x0, y0, i0;
        decode_m_tile(p, mt, x0, y0, i0);
        const int64_t m0 = ((int64_t)i0 * p.H + y0) * p.W + x0;  // tile rows are contiguous (checked on the host)
        const int64_t m = m0 + r;
        const bool row_ok = m < p.M;
        const int last_j = group + ((nslab - 1 - group) / 2) * 2;  // last slab this group handles
        // (1) bias row of this tile -> smem (host guarantees one bias row per tile: bias_div % 128 == 0)
        {
          const float* brow = p.bias + (p.bias_div > 0 ? (m0 / p.bias_div) * (int64_t)p.bias_ld : 0) + nt * BN;
          for (int c = gtid; c < BN; c += 128)
            my_bias[c] = (p.bias != nullptr && nt * BN + c < p.N) ? __ldg(brow + c) : 0.f;
        }
        // (2) residual fragments of every slab this group owns
        uint4 res[NSLAB_MAX][4];
        if (p.residual != nullptr) {
#pragma unroll
          for (int s = 0; s < NSLAB_MAX; ++s) {
            const int j = group + 2 * s;
            const int ncol0 = nt * BN + j * 32;
            if (j < nslab && row_ok && ncol0 + 32 <= n_out_total) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * (int64_t)p.ldr + ncol0);
#pragma unroll
              for (int e = 0; e < 4; ++e) res[s][e] = __ldg(rp + e);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) res[s][e] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        named_bar_sync(1 + group, 128);  // bias row visible to the group
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS;
        bool released = false;
#pragma unroll
        for (int s = 0; s < NSLAB_MAX; ++s) {
          const int j = group + 2 * s;
          if (j >= nslab) break;
          const int ncol0 = geglu ? nt * (BN / 2) + j * 32 : nt * BN + j * 32;  // first output column of the slab
          if (ncol0 >= n_out_total) break;  // this and all later slabs lie beyond N (uniform over the group)
          uint32_t v[32];
          float f[32];
          tmem_ld_32x32(tbase + j * 32, v);
          if (!geglu) {
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) {
                const float4 t = *reinterpret_cast<const float4*>(my_bias + j * 32 + e4 * 4);
                f[e4 * 4] += t.x;
                f[e4 * 4 + 1] += t.y;
                f[e4 * 4 + 2] += t.z;
                f[e4 * 4 + 3] += t.w;
              }
            }
          } else {
            uint32_t g[32];
            tmem_ld_32x32(tbase + BN / 2 + j * 32, g);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float fv = __uint_as_float(v[e]) + my_bias[j * 32 + e];
              const float fg = __uint_as_float(g[e]) + my_bias[BN / 2 + j * 32 + e];
              f[e] = fv * gelu_erf_f(fg);
            }
          }
          if (j == last_j) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            released = true;
          }
          if (p.residual != nullptr) {
            if (ncol0 + 32 <= n_out_total) {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&res[s][e4]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t = __half22float2(h2[e]);
                  f[e4 * 8 + e * 2] += t.x;
                  f[e4 * 8 + e * 2 + 1] += t.y;
                }
              }
            } else if (row_ok) {  // ragged last slab: scalar, in-bounds columns only
              const __half* rr = p.residual + m * (int64_t)p.ldr + ncol0;
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (ncol0 + e < n_out_total) f[e] += __half2float(rr[e]);
            }
          }
          if (p.flags & LS_EPI_SILU) {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = silu_f(f[e]);
          }
          // stage the slab: row r is 64 bytes, 16-byte chunk c lives at position c ^ ((r >> 1) & 3)  (SWIZZLE_64B)
          uint8_t* slab = my_stage + (slab_count & 1u) * SLAB_BYTES;
          ++slab_count;
          const int sw = (r >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[c * 8 + e * 2], f[c * 8 + e * 2 + 1]);
            *reinterpret_cast<uint4*>(slab + r * 64 + ((c ^ sw) << 4)) = u;
          }
          fence_proxy_async_smem();
          // every store issued before this slab has finished READING its smem => the other slab may be overwritten
          // once the group passes the barrier below
          if (issuer) bulk_wait_group_read<0>();
          named_bar_sync(1 + group, 128);
          if (issuer) {
            tma_store_2d(&p.mapOut, slab, ncol0, (int)m0);
            bulk_commit_group();
          }
        }
        if (!released) {  // group had no slab inside N for this tile: still hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (issuer) bulk_wait_group_read<0>();
    } else if (group == 0) {
    const int ix = r % p.bw;
    const int iy = (r / p.bw) % p.bh;
    const int in = r / (p.bw * p.bh);
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const uint32_t acc_phase = (lt >> 1) & 1u;
      const int mt = tile / p.n_tiles;
      const int nt = tile - mt * p.n_tiles;
      int x0, y0, i0;
      decode_m_tile(p, mt, x0, y0, i0);
      const int x = x0 + ix, y = y0 + iy, img = i0 + in;
      const bool row_ok = (x < p.W) && (y < p.H) && (img < p.nimg);
      const int64_t m = ((int64_t)img * p.H + y) * p.W + x;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS;

      if (!(p.flags & LS_EPI_GEGLU)) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + c * 32, v);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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
        if constexpr (BN >= 64) {
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c) {
            uint32_t v[32], g[32];
            tmem_ld_32x32(tbase + c * 32, v);
            tmem_ld_32x32(tbase + BN / 2 + c * 32, g);
            tmem_ld_wait();
            if (c == BN / 64 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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
    }
  } else {
      // legacy direct-store path uses the first 4 epilogue warps only; the others just release the accumulators
      int lt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1u;
        mbar_wait(&tmem_full[acc], acc_phase);
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
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
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int pick_tile_n(int m_tiles, int N, int sms) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  const int cands[5] = {256, 160, 128, 64, 32};
  int best = 128;
  long best_cost = -1;
  for (int i = 0; i < 5; ++i) {
    const int c = cands[i];
    const long n_tiles = (N + c - 1) / c;
    const long tiles = n_tiles * m_tiles;
    const long waves = (tiles + sms - 1) / sms;
    // per-tile time ~ BN MMA columns (+ fixed overhead); tiles narrower than 64 are operand-bandwidth bound
    const long cost = waves * ((c < 64 ? 64 : c) + 24);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

template <int BN>
static int launch_gemm(const GemmKParams& p, int grid, cudaStream_t stream) {
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BN * BK * 2;
  const size_t smem = (size_t)p.stages * STAGE_BYTES + STAGING_BYTES + 2048 + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    LS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, smem, stream>>>(p);
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

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
    LS_CHECK(a->a_taps[s] == 1 || a->a_taps[s] == 9, "ls_gemm: taps must be 1 or 9");
    LS_CHECK((reinterpret_cast<uintptr_t>(a->a_ptr[s]) & 15) == 0, "ls_gemm: segment %d pointer not 16B aligned", s);
    any_conv |= (a->a_taps[s] == 9);
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
  if (any_conv) {
    // a 3x3 tap shift must stay inside one image row/column block: the box spans whole rows (or W >= 128)
    LS_CHECK(a->b_batch_stride == 0, "ls_gemm: batched conv unsupported");
  }
  p.N = a->N;
  p.b_batched = a->b_batch_stride != 0 ? 1 : 0;
  if (p.b_batched) LS_CHECK(p.bn == 1, "ls_gemm: batched GEMM needs >= 128 rows per problem or H == 1");

  int BN = a->tile_n;
  const int sms = num_sms();
  if (a->flags & LS_EPI_GEGLU) {
    LS_CHECK(BN == 64 || BN == 128 || BN == 256, "ls_gemm: GEGLU needs explicit tile_n in {64,128,256}");
    LS_CHECK(a->N % BN == 0, "ls_gemm: GEGLU needs N %% tile_n == 0");
  }
  if (BN == 0) BN = pick_tile_n(p.m_tiles, a->N, sms);
  LS_CHECK(BN == 32 || BN == 64 || BN == 128 || BN == 160 || BN == 256, "ls_gemm: unsupported tile_n %d", BN);
  p.n_tiles = (a->N + BN - 1) / BN;

  // tensor maps
  for (int s = 0; s < a->nseg; ++s) {
    const cuuint64_t ld_b = (cuuint64_t)a->a_ld[s] * 2;
    cuuint64_t gdim[4] = {(cuuint64_t)a->a_ch[s], (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->nimg};
    cuuint64_t gstr[3] = {ld_b, ld_b * a->W, ld_b * a->W * a->H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
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
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(a->b_ptr), gdim, gstr, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }

  // output path: smem-staged TMA bulk stores when the tile's 128 rows are contiguous output rows and the row pitch
  // is 16-byte aligned; otherwise (fp32 output, ragged geometry, narrow ld) per-thread direct stores
  const int64_t M = (int64_t)a->nimg * a->H * a->W;
  p.M = M;
  const int n_out = (a->flags & LS_EPI_GEGLU) ? a->N / 2 : a->N;
  const bool rows_contig = (a->W < BM) || (a->W % BM == 0) || (a->H == 1 && a->nimg == 1);
  p.tma_store = (!(a->flags & LS_EPI_OUT_F32) && (a->ldo % 8 == 0) && rows_contig && M < (1ll << 31) &&
                 (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (a->bias_div == 0 || a->bias_div % BM == 0) &&
                 (a->residual == nullptr || ((a->ldr % 8 == 0) && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0)))
                    ? 1
                    : 0;
  if (p.tma_store) {
    cuuint64_t gdim[2] = {(cuuint64_t)n_out, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)a->ldo * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.mapOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->out, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LS_CHECK(r == CUDA_SUCCESS, "ls_gemm: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
  }
  p.bias = a->bias;
  p.bias_div = a->bias_div;
  p.bias_ld = a->bias_ld > 0 ? a->bias_ld : a->N;
  p.residual = reinterpret_cast<const __half*>(a->residual);
  p.ldr = a->ldr;
  p.out = a->out;
  p.ldo = a->ldo;
  p.flags = a->flags;

  const int stage_bytes = A_STAGE_BYTES + BN * BK * 2;
  int stages = (227 * 1024 - 1024 - 256 - 2048 - STAGING_BYTES) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > p.num_kb && p.num_kb >= 2) stages = p.num_kb;
  if (stages < 2) stages = 2;
  p.stages = stages;

  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sms ? total : sms;
  switch (BN) {
    case 32: return launch_gemm<32>(p, grid, stream);
    case 64: return launch_gemm<64>(p, grid, stream);
    case 128: return launch_gemm<128>(p, grid, stream);
    case 160: return launch_gemm<160>(p, grid, stream);
    case 256: return launch_gemm<256>(p, grid, stream);
  }
  return 1;
}

}  // namespace ls

extern "C" int ls_gemm(const LsGemmArgs* args, void* stream) {
  return ls::gemm_impl(args, reinterpret_cast<cudaStream_t>(stream));
}
