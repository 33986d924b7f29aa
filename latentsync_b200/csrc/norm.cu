// GroupNorm (two-phase: fp32 statistics + fused normalise/affine/SiLU), LayerNorm (+ temporal positional table),
// row softmax and transpose.  All memory-bound: 16-byte vectorised, coalesced, warp-shuffle reductions.
//
// Reference: nn.GroupNorm at latentsync/models/resnet.py:140,164 (5-D input => statistics span all 16 frames),
// attention.py:51 and motion_module.py:101 (per frame), unet.py:236; nn.LayerNorm at attention.py:145,157,172 and
// motion_module.py:195,201; PositionalEncoding motion_module.py:221-234.
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>

namespace ls {

extern std::atomic<int64_t> g_launch_count;

// -------------------------------------------------------------------------------------------------- GroupNorm
// Deterministic two-level reduction (no floating-point atomics): every CTA reduces its row chunk with 16-byte loads
// (thread = one 8-channel column vector x one row lane, several rows in flight), folds row lanes and channels into groups
// in a fixed order and writes one partial per (instance, chunk, group); the CTA that arrives last at the per-instance
// ticket (an integer atomic) sums the partials in chunk order and resets the ticket.
__global__ void __launch_bounds__(256) gn_stats_kernel(const __half* __restrict__ x1, int c1,
                                                       const __half* __restrict__ x2, int c2, int rows_per_inst,
                                                       int rows_per_chunk, int groups, float* __restrict__ partial,
                                                       float* __restrict__ stats, unsigned int* __restrict__ tickets) {
  pdl_prologue();
  extern __shared__ float gs_sm[];  // [RL][C] sums, then [RL][C] sums of squares
  __shared__ int s_last;
  const int C = c1 + c2;
  const int cg = C / groups;
  const int nvec = C >> 3;
  const int RL = (nvec <= (int)blockDim.x) ? (int)blockDim.x / nvec : 1;  // row lanes
  float* sh_s = gs_sm;
  float* sh_q = gs_sm + RL * C;
  const int inst = blockIdx.y;
  const int chunks = gridDim.x;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)blockIdx.x * rows_per_chunk;
  int64_t row_end = row0 + rows_per_chunk;
  const int64_t inst_end = (int64_t)(inst + 1) * rows_per_inst;
  if (row_end > inst_end) row_end = inst_end;
  const int rl = (nvec <= (int)blockDim.x) ? (int)threadIdx.x / nvec : 0;
  const int cv0 = (nvec <= (int)blockDim.x) ? (int)threadIdx.x % nvec : (int)threadIdx.x;
  const int cvstep = (nvec <= (int)blockDim.x) ? nvec : (int)blockDim.x;
  if (rl < RL) {
    for (int cv = cv0; cv < nvec; cv += cvstep) {
      const int c = cv * 8;
      const __half* src = (c < c1) ? (x1 + c) : (x2 + (c - c1));
      const int ld = (c < c1) ? c1 : c2;
      float s[8], q[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
      int64_t row = row0 + rl;
      for (; row + 3 * RL < row_end; row += 4 * RL) {  // four rows in flight
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            s[2 * e] += f.x;
            s[2 * e + 1] += f.y;
            q[2 * e] += f.x * f.x;
            q[2 * e + 1] += f.y * f.y;
          }
        }
      }
      for (; row < row_end; row += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + row * ld));
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          s[2 * e] += f.x;
          s[2 * e + 1] += f.y;
          q[2 * e] += f.x * f.x;
          q[2 * e + 1] += f.y * f.y;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sh_s[rl * C + c + e] = s[e];
        sh_q[rl * C + c + e] = q[e];
      }
    }
  }
  __syncthreads();
  float* my_partial = partial + ((int64_t)inst * chunks + blockIdx.x) * groups * 2;
  if (threadIdx.x < groups * 2) {  // thread = (group, stat); fixed summation order
    const int g = threadIdx.x >> 1;
    const float* base = (threadIdx.x & 1) ? sh_q : sh_s;
    float acc = 0.f;
    for (int r = 0; r < RL; ++r)
      for (int i = 0; i < cg; ++i) acc += base[r * C + g * cg + i];
    my_partial[threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(&tickets[inst], 1u);
    s_last = (old == (unsigned int)chunks - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // 4 threads per (group, stat): interleaved chunk ranges, combined in a fixed shuffle order
  const int item = threadIdx.x >> 2, part = threadIdx.x & 3;
  const bool active = item < groups * 2;
  float acc = 0.f;
  if (active) {
    const float* base = partial + (int64_t)inst * chunks * groups * 2 + item;
    for (int ch = part; ch < chunks; ch += 4) acc += __ldcg(base + (int64_t)ch * groups * 2);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (active && part == 0) stats[(int64_t)inst * groups * 2 + item] = acc;
  if (threadIdx.x == 0) tickets[inst] = 0u;
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x1, int c1,
                                                       const __half* __restrict__ x2, int c2, int rows_per_inst,
                                                       int rows_per_chunk, int groups, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ y) {
  pdl_prologue();
  extern __shared__ float gn_sm[];
  const int C = c1 + c2;
  float* sa = gn_sm;
  float* sb = gn_sm + C;
  const int cg = C / groups;
  const int inst = blockIdx.y;
  const float inv_n = 1.f / ((float)rows_per_inst * (float)cg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cg;
    const float mean = stats[((int64_t)inst * groups + g) * 2] * inv_n;
    float var = stats[((int64_t)inst * groups + g) * 2 + 1] * inv_n - mean * mean;
    var = fmaxf(var, 0.f);
    const float a = rsqrtf(var + eps) * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)blockIdx.x * rows_per_chunk;
  int64_t row_end = row0 + rows_per_chunk;
  const int64_t inst_end = (int64_t)(inst + 1) * rows_per_inst;
  if (row_end > inst_end) row_end = inst_end;
  const int64_t total = (row_end - row0) * nvec;
  for (int64_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int64_t row = row0 + idx / nvec;
    const int c = (int)(idx % nvec) * 8;
    const __half* src = (c < c1) ? (x1 + row * c1 + c) : (x2 + row * c2 + (c - c1));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
    uint4 w;
    __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h2[e]);
      float a = f.x * sa[c + 2 * e] + sb[c + 2 * e];
      float b = f.y * sa[c + 2 * e + 1] + sb[c + 2 * e + 1];
      if (silu) {
        a = silu_f(a);
        b = silu_f(b);
      }
      o2[e] = __floats2half2_rn(a, b);
    }
    *reinterpret_cast<uint4*>(y + row * C + c) = w;
  }
}

// Fused GroupNorm: statistics and normalisation in ONE launch.  Every CTA reduces its row chunk (same code path as
// gn_stats_kernel), publishes the partial and meets the other CTAs of its instance at a ticket; after the rendezvous
// each CTA sums the instance's partials in chunk order (deterministic, redundantly: chunks x 64 floats from L2),
// builds the per-channel scale/shift in smem and normalises the chunk it has just read (L2-hot).  The host guarantees
// that the whole grid is co-resident (grid <= occupancy x SMs, stream-ordered launches), so the rendezvous cannot hang.
__global__ void __launch_bounds__(256, 3) gn_fused_kernel(const __half* __restrict__ x1, int c1,
                                                       const __half* __restrict__ x2, int c2, int rows_per_inst,
                                                       int rows_per_chunk, int groups, float* __restrict__ partial,
                                                       unsigned int* __restrict__ tickets,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ y) {
  pdl_prologue();
  extern __shared__ float gf_sm[];  // phase 1: [RL][C] sums + [RL][C] squares; phase 2: sa[C], sb[C] (aliased)
  __shared__ float s_stats[64];
  const int C = c1 + c2;
  const int cg = C / groups;
  const int nvec = C >> 3;
  const int RL = (nvec <= (int)blockDim.x) ? (int)blockDim.x / nvec : 1;
  float* sh_s = gf_sm;
  float* sh_q = gf_sm + RL * C;
  const int inst = blockIdx.y;
  const int chunks = gridDim.x;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)blockIdx.x * rows_per_chunk;
  int64_t row_end = row0 + rows_per_chunk;
  const int64_t inst_end = (int64_t)(inst + 1) * rows_per_inst;
  if (row_end > inst_end) row_end = inst_end;
  const int rl = (nvec <= (int)blockDim.x) ? (int)threadIdx.x / nvec : 0;
  const int cv0 = (nvec <= (int)blockDim.x) ? (int)threadIdx.x % nvec : (int)threadIdx.x;
  const int cvstep = (nvec <= (int)blockDim.x) ? nvec : (int)blockDim.x;
  if (rl < RL) {
    for (int cv = cv0; cv < nvec; cv += cvstep) {
      const int c = cv * 8;
      const __half* src = (c < c1) ? (x1 + c) : (x2 + (c - c1));
      const int ld = (c < c1) ? c1 : c2;
      float s[8], q[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
      int64_t row = row0 + rl;
      for (; row + 7 * RL < row_end; row += 8 * RL) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            s[2 * e] += f.x;
            s[2 * e + 1] += f.y;
            q[2 * e] += f.x * f.x;
            q[2 * e + 1] += f.y * f.y;
          }
        }
      }
      for (; row < row_end; row += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + row * ld));
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          s[2 * e] += f.x;
          s[2 * e + 1] += f.y;
          q[2 * e] += f.x * f.x;
          q[2 * e + 1] += f.y * f.y;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sh_s[rl * C + c + e] = s[e];
        sh_q[rl * C + c + e] = q[e];
      }
    }
  }
  __syncthreads();
  float* inst_partial = partial + (int64_t)inst * chunks * groups * 2;
  if (threadIdx.x < groups * 2) {
    const int g = threadIdx.x >> 1;
    const float* base = (threadIdx.x & 1) ? sh_q : sh_s;
    float acc = 0.f;
    for (int r = 0; r < RL; ++r)
      for (int i = 0; i < cg; ++i) acc += base[r * C + g * cg + i];
    inst_partial[(int64_t)blockIdx.x * groups * 2 + threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  unsigned int* tk = tickets + 2 * inst;
  // affine parameters of this thread's first column vector: requested before the rendezvous (latency hides behind it)
  float4 gpre[2], bpre[2];
  {
    const int c = (cv0 < nvec ? cv0 : 0) * 8;
    gpre[0] = __ldg(reinterpret_cast<const float4*>(gamma + c));
    gpre[1] = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    bpre[0] = __ldg(reinterpret_cast<const float4*>(beta + c));
    bpre[1] = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
  }
  if (threadIdx.x == 0) {
    atomicAdd(tk, 1u);
    uint32_t spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(tk) < (unsigned int)chunks) {
      __nanosleep(64);
      if (++spins > (1u << 24)) {
        printf("latentsync_b200: GroupNorm rendezvous timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
        __trap();
      }
    }
  }
  __syncthreads();
  __threadfence();
  {  // 4 threads per (group, stat): interleaved chunk ranges, combined in a fixed shuffle order.  16 independent loads
    // per round: with a single running sum the 74 L2 reads per thread of a 296-chunk instance went out a few at a time
    // (ncu: joint-statistics launches 12 us slower than per-frame ones of the same tensor).
    const int item = threadIdx.x >> 2, part = threadIdx.x & 3;
    const bool active = item < groups * 2;
    float acc = 0.f;
    if (active) {
      float a16[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a16[k] = 0.f;
      const float* pp = inst_partial + item;
      const int64_t ps = (int64_t)groups * 2;
      int ch = part;
      for (; ch + 60 < chunks; ch += 64) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a16[k] += __ldcg(pp + (int64_t)(ch + 4 * k) * ps);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (ch + 4 * k < chunks) a16[k] += __ldcg(pp + (int64_t)(ch + 4 * k) * ps);
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) a16[k] += a16[k + w];
      acc = a16[0];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (active && part == 0) s_stats[item] = acc;
  }
  __syncthreads();
  // Normalise the chunk just read (L2-hot).  Each thread keeps the column vector it reduced (8 channels): scale / shift
  // live in 16 registers (per-channel smem tables read with stride 32 B were 8-way bank conflicted and made this phase
  // MIO bound), rows advance RL at a time with four independent 16-byte loads in flight.
  const float inv_n = 1.f / ((float)rows_per_inst * (float)cg);
  if (rl < RL) {
    for (int cv = cv0; cv < nvec; cv += cvstep) {
      const int c = cv * 8;
      float ga[8], be[8];
      if (cv == cv0) {
        ga[0] = gpre[0].x; ga[1] = gpre[0].y; ga[2] = gpre[0].z; ga[3] = gpre[0].w;
        ga[4] = gpre[1].x; ga[5] = gpre[1].y; ga[6] = gpre[1].z; ga[7] = gpre[1].w;
        be[0] = bpre[0].x; be[1] = bpre[0].y; be[2] = bpre[0].z; be[3] = bpre[0].w;
        be[4] = bpre[1].x; be[5] = bpre[1].y; be[6] = bpre[1].z; be[7] = bpre[1].w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ga[e] = __ldg(gamma + c + e);
          be[e] = __ldg(beta + c + e);
        }
      }
      float sa8[8], sb8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = (c + e) / cg;
        const float mean = s_stats[2 * g] * inv_n;
        float var = s_stats[2 * g + 1] * inv_n - mean * mean;
        var = fmaxf(var, 0.f);
        sa8[e] = rsqrtf(var + eps) * ga[e];
        sb8[e] = be[e] - mean * sa8[e];
      }
      const __half* src = (c < c1) ? (x1 + c) : (x2 + (c - c1));
      const int ld = (c < c1) ? c1 : c2;
      auto apply_store = [&](const uint4& u, int64_t row) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        uint4 w;
        __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          float a = f.x * sa8[2 * e] + sb8[2 * e];
          float b = f.y * sa8[2 * e + 1] + sb8[2 * e + 1];
          if (silu) {
            a = silu_f(a);
            b = silu_f(b);
          }
          o2[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(y + row * C + c) = w;
      };
      int64_t row = row0 + rl;
      for (; row + 7 * RL < row_end; row += 8 * RL) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 8; ++k) apply_store(u[k], row + k * RL);
      }
      for (; row + 1 * RL < row_end; row += 2 * RL) {
        uint4 u[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 2; ++k) apply_store(u[k], row + k * RL);
      }
      for (; row < row_end; row += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + row * ld));
        apply_store(u, row);
      }
    }
  }
  // leave: the last CTA of the instance resets both counters for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(tk + 1, 1u);
    if (old == (unsigned int)chunks - 1u) {
      tk[0] = 0u;
      tk[1] = 0u;
    }
  }
}


// -------------------------------------------------------------------------------------------------------------
// GroupNorm from the producer's partials: the GEMM that wrote x also wrote, per 128-row tile and per `unit` columns, the
// (sum, sum of squares) of what it stored (LsGemmArgs.gn_partials_out).  No statistics pass over the tensor and no
// cross-CTA rendezvous: every CTA sums the few partials of its instance's groups (fixed order: deterministic), then makes
// ONE read-modify-write pass over its rows.  grid = (chunks * csplit, instances): a CTA owns a row chunk and 1 / csplit of
// the channels (whole groups), so that wide concatenations do not make every CTA read every partial.
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) gn_parts_kernel(const __half* __restrict__ x1, int c1, const float2* __restrict__ p1,
                                                          int ld1, const __half* __restrict__ x2, int c2,
                                                          const float2* __restrict__ p2, int ld2, int rows_per_inst,
                                                          int rows_per_chunk, int groups, int unit, int csplit,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, int silu, __half* __restrict__ y) {
  pdl_prologue();
  __shared__ float s_stats[64];
  const int C = c1 + c2;
  const int cg = C / groups;
  const int upg = cg / unit;       // units per group
  const int cs = (int)blockIdx.x % csplit;
  const int chunk = (int)blockIdx.x / csplit;
  const int Cl = C / csplit;       // this CTA's channels [cbeg, cbeg + Cl): whole groups, whole 8-channel vectors
  const int cbeg = cs * Cl;
  const int g0 = cbeg / cg, ng = Cl / cg;
  const int inst = blockIdx.y;
  const int tpi = rows_per_inst >> 7;  // 128-row tiles per instance
  const int u1 = c1 / unit;            // units of the first source
  // (1) statistics: tpg threads per group (a power of two, all of them inside one warp) take the group's tiles
  //     part, part + tpg, ... x its units into eight interleaved chains - every load of the CTA is in flight at once -
  //     then chains and threads are folded in a fixed order
  {
    int tpg = 32;
    while (tpg > 1 && tpg * ng > (int)blockDim.x) tpg >>= 1;
    const int gi = (int)threadIdx.x / tpg, part = (int)threadIdx.x % tpg;
    float su[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (gi < ng) {
      const int ug0 = (g0 + gi) * upg;
      const float2* b1 = p1 + (int64_t)inst * tpi * ld1;
      const float2* b2 = p2 + (int64_t)inst * tpi * ld2 - u1;
      // unit by unit: chain k takes the unit's tiles part + (8 i + k) tpg - eight independent loads per trip, static
      // register indices, 32-bit offsets from one base pointer per unit
      const int cnt = (tpi - part + tpg - 1) / tpg;  // tiles of this thread
      for (int ul = 0; ul < upg; ++ul) {
        const int ug = ug0 + ul;
        const bool first = ug < u1;
        const int ld = first ? ld1 : ld2;
        const float2* q = (first ? b1 + ug : b2 + ug) + part * ld;  // b2 is already offset by -u1
        const int step = tpg * ld;
        int i = 0;
        for (; i + 8 <= cnt; i += 8) {
          float2 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __ldg(q + (i + k) * step);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            su[k] += v[k].x;
            sq[k] += v[k].y;
          }
        }
        for (; i < cnt; ++i) {
          const float2 v = __ldg(q + i * step);
          su[0] += v.x;
          sq[0] += v.y;
        }
      }
    }
    float a = ((su[0] + su[1]) + (su[2] + su[3])) + ((su[4] + su[5]) + (su[6] + su[7]));
    float b = ((sq[0] + sq[1]) + (sq[2] + sq[3])) + ((sq[4] + sq[5]) + (sq[6] + sq[7]));
    for (int o = tpg >> 1; o >= 1; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (gi < ng && part == 0) {
      s_stats[2 * gi] = a;
      s_stats[2 * gi + 1] = b;
    }
  }
  __syncthreads();
  // (2) one pass: thread = (row lane, 8-channel vector), scale / shift of its vector in registers
  const int nvec = Cl >> 3;
  const int RL = (nvec <= (int)blockDim.x) ? (int)blockDim.x / nvec : 1;
  const int rl = (nvec <= (int)blockDim.x) ? (int)threadIdx.x / nvec : 0;
  const int cv0 = (nvec <= (int)blockDim.x) ? (int)threadIdx.x % nvec : (int)threadIdx.x;
  const int cvstep = (nvec <= (int)blockDim.x) ? nvec : (int)blockDim.x;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)chunk * rows_per_chunk;
  int64_t row_end = row0 + rows_per_chunk;
  const int64_t inst_end = (int64_t)(inst + 1) * rows_per_inst;
  if (row_end > inst_end) row_end = inst_end;
  const float inv_n = 1.f / ((float)rows_per_inst * (float)cg);
  if (rl < RL) {
    for (int cv = cv0; cv < nvec; cv += cvstep) {
      const int c = cbeg + cv * 8;
      float sa8[8], sb8[8];
      {
        const float4 ga0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 ga1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 be0 = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 be1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        const float ga[8] = {ga0.x, ga0.y, ga0.z, ga0.w, ga1.x, ga1.y, ga1.z, ga1.w};
        const float be[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int gi = (c + e) / cg - g0;
          const float mean = s_stats[2 * gi] * inv_n;
          float var = s_stats[2 * gi + 1] * inv_n - mean * mean;
          var = fmaxf(var, 0.f);
          sa8[e] = rsqrtf(var + eps) * ga[e];
          sb8[e] = be[e] - mean * sa8[e];
        }
      }
      const __half* src = (c < c1) ? (x1 + c) : (x2 + (c - c1));
      const int ld = (c < c1) ? c1 : c2;
      auto apply_store = [&](const uint4& u, int64_t row) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        uint4 w;
        __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          float a = f.x * sa8[2 * e] + sb8[2 * e];
          float b = f.y * sa8[2 * e + 1] + sb8[2 * e + 1];
          if (silu) {
            a = silu_f(a);
            b = silu_f(b);
          }
          o2[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(y + row * C + c) = w;
      };
      int64_t row = row0 + rl;
      for (; row + 7 * RL < row_end; row += 8 * RL) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 8; ++k) apply_store(u[k], row + k * RL);
      }
      for (; row + 1 * RL < row_end; row += 2 * RL) {
        uint4 u[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (row + k * RL) * ld));
#pragma unroll
        for (int k = 0; k < 2; ++k) apply_store(u[k], row + k * RL);
      }
      for (; row < row_end; row += RL) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + row * ld));
        apply_store(u, row);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// Cluster GroupNorm: one thread-block CLUSTER per instance.  Every CTA copies its rows ONCE into shared memory
// (cp.async, the whole chunk in flight), reduces them there, the CTAs exchange their 64 (group, stat) partials through
// distributed shared memory behind a hardware cluster barrier, and each CTA normalises its rows from shared memory.
// DRAM/L2 traffic is one read + one write of the tensor, and the rendezvous costs a cluster barrier instead of the
// global-memory ticket + fence + partial round trips of gn_fused_kernel (which needs ~11 us even for a 1 MB tensor).
// Used whenever an instance fits the shared memory of <= 16 CTAs: all per-frame norms of the transformer / motion
// modules, the joint norms of the 8x8 and 4x4 levels, the VAE mid block.  Sums in fixed rank order: deterministic.
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gn_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void gn_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float gn_ld_remote(const float* local_ptr, uint32_t rank) {
  uint32_t raddr;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
}

constexpr int GNC_THREADS = 512;

__global__ void __launch_bounds__(GNC_THREADS) gn_cluster_kernel(const __half* __restrict__ x1, int c1,
                                                                 const __half* __restrict__ x2, int c2,
                                                                 int rows_per_inst, int rows_per_cta, int cl, int groups,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, int silu,
                                                                 __half* __restrict__ y) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t gnc_sm[];
  __shared__ float s_part[64];   // this CTA's (group, stat) partial sums, read by the whole cluster
  __shared__ float s_stats[64];  // instance totals
  const int C = c1 + c2;
  const int cg = C / groups;
  const int nvec = C >> 3;
  const int tid = threadIdx.x;
  const uint32_t rank = gn_cluster_rank();
  const int inst = blockIdx.x / cl;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)rank * rows_per_cta;
  __half* data = reinterpret_cast<__half*>(gnc_sm);                                   // [rows_per_cta][C]
  float* red = reinterpret_cast<float*>(gnc_sm + (size_t)rows_per_cta * C * 2);       // [2][RL][C], later sa[C], sb[C]
  const int RL = (nvec <= GNC_THREADS) ? GNC_THREADS / nvec : 1;  // row lanes: thread = (row lane, column vector)
  const int rl = (nvec <= GNC_THREADS) ? tid / nvec : 0;
  const int cv0 = (nvec <= GNC_THREADS) ? tid % nvec : tid;
  const int cvstep = (nvec <= GNC_THREADS) ? nvec : GNC_THREADS;
  float* sh_s = red;
  float* sh_q = red + RL * C;

  // affine parameters of this thread's column vector first: their latency hides behind the copy
  float4 gpre[2], bpre[2];
  {
    const int c = (cv0 < nvec ? cv0 : 0) * 8;
    gpre[0] = __ldg(reinterpret_cast<const float4*>(gamma + c));
    gpre[1] = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    bpre[0] = __ldg(reinterpret_cast<const float4*>(beta + c));
    bpre[1] = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
  }
  // phase 0: the whole chunk global -> shared, every 16-byte vector requested at once
  const int total = rows_per_cta * nvec;
  for (int idx = tid; idx < total; idx += GNC_THREADS) {
    const int r = idx / nvec;
    const int c = (idx - r * nvec) * 8;
    const __half* src = (c < c1) ? (x1 + (row0 + r) * c1 + c) : (x2 + (row0 + r) * c2 + (c - c1));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(data + (size_t)r * C + c)), "l"(src)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // phase 1: per-channel sums over the chunk (from shared memory), then per-group
  {
    if (rl < RL) {
      for (int cv = cv0; cv < nvec; cv += cvstep) {
        float sacc[8], qacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) sacc[e] = qacc[e] = 0.f;
        for (int r = rl; r < rows_per_cta; r += RL) {
          const uint4 u = *reinterpret_cast<const uint4*>(data + (size_t)r * C + cv * 8);
          const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            sacc[2 * e] += f.x;
            sacc[2 * e + 1] += f.y;
            qacc[2 * e] += f.x * f.x;
            qacc[2 * e + 1] += f.y * f.y;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sh_s[rl * C + cv * 8 + e] = sacc[e];
          sh_q[rl * C + cv * 8 + e] = qacc[e];
        }
      }
    }
  }
  __syncthreads();
  {  // 8 threads per (group, stat): fixed interleave + fixed shuffle order
    const int item = tid >> 3, part = tid & 7;
    float acc = 0.f;
    if (item < groups * 2) {
      const int g = item >> 1;
      const float* base = (item & 1) ? sh_q : sh_s;
      for (int i = part; i < RL * cg; i += 8) {
        const int r = i / cg;
        acc += base[r * C + g * cg + (i - r * cg)];
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (item < 64 && part == 0) s_part[item] = (item < groups * 2) ? acc : 0.f;
  }
  // rendezvous of the cluster, then every CTA sums the partials of all ranks in rank order
  gn_cluster_arrive();
  gn_cluster_wait();
  if (tid < 64) {
    float acc = 0.f;
    for (int r = 0; r < cl; ++r) acc += gn_ld_remote(&s_part[tid], (uint32_t)r);
    s_stats[tid] = acc;
  }
  gn_cluster_arrive();  // this CTA no longer reads remote shared memory (waited for at the very end)
  __syncthreads();
  // phase 2: normalise from shared memory; the thread keeps its column vector, scale / shift in 16 registers
  const float inv_n = 1.f / ((float)rows_per_inst * (float)cg);
  if (rl < RL) {
    for (int cv = cv0; cv < nvec; cv += cvstep) {
      const int c = cv * 8;
      float ga[8], be[8];
      if (cv == cv0) {
        ga[0] = gpre[0].x; ga[1] = gpre[0].y; ga[2] = gpre[0].z; ga[3] = gpre[0].w;
        ga[4] = gpre[1].x; ga[5] = gpre[1].y; ga[6] = gpre[1].z; ga[7] = gpre[1].w;
        be[0] = bpre[0].x; be[1] = bpre[0].y; be[2] = bpre[0].z; be[3] = bpre[0].w;
        be[4] = bpre[1].x; be[5] = bpre[1].y; be[6] = bpre[1].z; be[7] = bpre[1].w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ga[e] = __ldg(gamma + c + e);
          be[e] = __ldg(beta + c + e);
        }
      }
      float sa8[8], sb8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = (c + e) / cg;
        const float mean = s_stats[2 * g] * inv_n;
        float var = s_stats[2 * g + 1] * inv_n - mean * mean;
        var = fmaxf(var, 0.f);
        sa8[e] = rsqrtf(var + eps) * ga[e];
        sb8[e] = be[e] - mean * sa8[e];
      }
      for (int r = rl; r < rows_per_cta; r += RL) {
        const uint4 u = *reinterpret_cast<const uint4*>(data + (size_t)r * C + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        uint4 w;
        __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          float a = f.x * sa8[2 * e] + sb8[2 * e];
          float b = f.y * sa8[2 * e + 1] + sb8[2 * e + 1];
          if (silu) {
            a = silu_f(a);
            b = silu_f(b);
          }
          o2[e] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(y + (row0 + r) * C + c) = w;
      }
    }
  }
  gn_cluster_wait();  // nobody reads this CTA's s_part any more: safe to exit
}

// -------------------------------------------------------------------------------------------------------------
// Group-local GroupNorm for SMALL instances: one CTA per (instance, group).  The slab of one group - rows_per_inst rows
// x C / groups channels - is a few KB at the 8x8 / 4x4 levels (256 rows x 40 channels = 20 KB), so one CTA reads it
// (pass 1: sums), reduces inside the block and reads it again from L1 / L2 to normalise (pass 2).  No cross-CTA
// exchange at all: the cluster kernel above pays a cluster barrier + DSMEM round trip (~8-12 us per launch on tensors of
// 1-5 MB); this one is bound by two dependent load round trips.  Fixed summation order: deterministic.
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_group_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2,
                                                       int c2, int rows_per_inst, int groups,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ y) {
  pdl_prologue();
  __shared__ float red[2][8];
  __shared__ float s_mean, s_rstd;
  const int C = c1 + c2;
  const int cg = C / groups;
  const int g = blockIdx.x, inst = blockIdx.y;
  const int nvec = cg >> 2;  // 8-byte vectors (4 channels) per row
  const int c0 = g * cg;
  const int64_t row0 = (int64_t)inst * rows_per_inst;
  const int total = rows_per_inst * nvec;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto src_of = [&](int row, int v) -> const __half* {
    const int c = c0 + v * 4;
    return (c < c1) ? x1 + (row0 + row) * c1 + c : x2 + (row0 + row) * c2 + (c - c1);
  };
  float su = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / nvec, v = i - row * nvec;
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src_of(row, v)));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    su += (a.x + a.y) + (b.x + b.y);
    sq = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(b.x, b.x, fmaf(b.y, b.y, sq))));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) {
    red[0][warp] = su;
    red[1][warp] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a += red[0][k];
      b += red[1][k];
    }
    const float inv_n = 1.f / ((float)rows_per_inst * (float)cg);
    const float mean = a * inv_n;
    const float var = fmaxf(b * inv_n - mean * mean, 0.f);
    s_mean = mean;
    s_rstd = rsqrtf(var + eps);
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / nvec, v = i - row * nvec;
    const int c = c0 + v * 4;
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(src_of(row, v)));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    float o0 = (a.x - mean) * rstd * ga.x + be.x, o1 = (a.y - mean) * rstd * ga.y + be.y;
    float o2 = (b.x - mean) * rstd * ga.z + be.z, o3 = (b.y - mean) * rstd * ga.w + be.w;
    if (silu) {
      o0 = silu_f(o0);
      o1 = silu_f(o1);
      o2 = silu_f(o2);
      o3 = silu_f(o3);
    }
    uint2 w;
    *reinterpret_cast<__half2*>(&w.x) = __floats2half2_rn(o0, o1);
    *reinterpret_cast<__half2*>(&w.y) = __floats2half2_rn(o2, o3);
    *reinterpret_cast<uint2*>(y + (row0 + row) * C + c) = w;
  }
}

static void gn_chunking(int64_t rows, int rows_per_inst, int target_ctas, int& ninst, int& chunks, int& rpc) {
  ninst = (int)(rows / rows_per_inst);
  chunks = (target_ctas + ninst - 1) / ninst;
  if (chunks < 1) chunks = 1;
  int max_chunks = rows_per_inst / 8;
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  rpc = (rows_per_inst + chunks - 1) / chunks;
  chunks = (rows_per_inst + rpc - 1) / rpc;
}

// --------------------------------------------------------------------------------------------------- LayerNorm
// One warp normalises ROWS rows at a time (all their 16-byte loads are issued before the first reduction, which is
// what keeps enough bytes in flight per SM: one row per warp ran at 2.2 TB/s), values stay in registers, exact
// two-pass variance.  VPL = 16-byte vectors per lane: C <= 256 * VPL.
template <int VPL, int ROWS>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int64_t rows, int C,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        const float* __restrict__ pe, int rows_per_frame, int nframes,
                                                        __half* __restrict__ y) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= rows) return;
  const int nvec = C >> 3;
  uint4 raw[ROWS][VPL];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int64_t row = row0 + r;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      raw[r][i] = (row < rows && vi < nvec) ? __ldg(reinterpret_cast<const uint4*>(x + row * C + vi * 8))
                                            : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int64_t row = row0 + r;
    if (row >= rows) break;  // warp-uniform
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw[r][i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        sum += f.x + f.y;  // padded vectors are zero
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[i][e] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    const float* pe_row = nullptr;
    if (pe != nullptr) pe_row = pe + (int64_t)((row / rows_per_frame) % nframes) * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const int c = vi * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float rr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) rr[e] = (v[i][e] - mean) * rstd * gg[e] + bb[e];
        if (pe_row != nullptr) {
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(pe_row + c));
          const float4 p1 = __ldg(reinterpret_cast<const float4*>(pe_row + c + 4));
          rr[0] += p0.x; rr[1] += p0.y; rr[2] += p0.z; rr[3] += p0.w;
          rr[4] += p1.x; rr[5] += p1.y; rr[6] += p1.z; rr[7] += p1.w;
        }
        uint4 w;
        __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(rr[2 * e], rr[2 * e + 1]);
        *reinterpret_cast<uint4*>(y + row * C + c) = w;
      }
    }
  }
}

// C = 40 * LPR (320 / 640 / 1280: every width of the stage2 UNet): LPR lanes share a row, each lane owns FIVE 16-byte
// vectors (vector index sub + LPR * i: a row's lanes read 128-byte-contiguous runs), so every lane of the warp carries
// data.  With one row per warp and 32-lane strides (kernel above) a 320-wide row keeps 40 of 64 vector slots busy and
// the warp still issues both slots' instructions: ncu counted 26 thread instructions per element, IPC 0.55 at 31 %
// occupancy, 17.8 us for the 21 MB level-0 tensor.  P row groups per warp keep 5 * P 16-byte loads in flight per lane.
template <int LPR, int P>
__global__ void __launch_bounds__(256) layernorm40_kernel(const __half* __restrict__ x, int64_t rows, int C,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          const float* __restrict__ pe, int rows_per_frame, int nframes,
                                                          __half* __restrict__ y) {
  pdl_prologue();
  constexpr int RPP = 32 / LPR;  // rows per pass of the warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsub = lane / LPR;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (RPP * P);
  if (row0 >= rows) return;
  uint4 raw[P][5];
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int64_t row = row0 + q * RPP + rsub;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      raw[q][i] = (row < rows) ? __ldg(reinterpret_cast<const uint4*>(x + row * C + (sub + LPR * i) * 8))
                               : make_uint4(0u, 0u, 0u, 0u);
  }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int64_t row = row0 + q * RPP + rsub;
    float v[5][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw[q][i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        sum += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq = fmaf(d, d, sq);
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    if (row >= rows) continue;
    const float* pe_row = nullptr;
    if (pe != nullptr) pe_row = pe + (int64_t)((row / rows_per_frame) % nframes) * C;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int c = (sub + LPR * i) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float rr[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) rr[e] = (v[i][e] - mean) * rstd * gg[e] + bb[e];
      if (pe_row != nullptr) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pe_row + c));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pe_row + c + 4));
        rr[0] += p0.x; rr[1] += p0.y; rr[2] += p0.z; rr[3] += p0.w;
        rr[4] += p1.x; rr[5] += p1.y; rr[6] += p1.z; rr[7] += p1.w;
      }
      uint4 w;
      __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(rr[2 * e], rr[2 * e + 1]);
      *reinterpret_cast<uint4*>(y + row * C + c) = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------- row softmax
constexpr int SM_MAXV = 8;  // cols <= 8 * 32 * 8 = 2048

__global__ void softmax_rows_kernel(const float* __restrict__ s, int64_t rows, int cols, float scale,
                                    __half* __restrict__ p) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 3;
  const float sl = scale * 1.4426950408889634f;
  float v[SM_MAXV][8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const float4 a = *reinterpret_cast<const float4*>(s + row * cols + vi * 8);
      const float4 b = *reinterpret_cast<const float4*>(s + row * cols + vi * 8 + 4);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[i][e]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[i][e] = mufu_ex2((v[i][e] - mx) * sl);
        sum += v[i][e];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      uint4 w;
      __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(v[i][2 * e] * inv, v[i][2 * e + 1] * inv);
      *reinterpret_cast<uint4*>(p + row * cols + vi * 8) = w;
    }
  }
}

// --------------------------------------------------------------------------------------------------- transpose
__global__ void transpose_kernel(const __half* __restrict__ x, int R, int C, __half* __restrict__ y) {
  pdl_prologue();
  __shared__ __half tile[32][34];
  const int64_t boff = (int64_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = x[boff + (int64_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) y[boff + (int64_t)c * R + r] = tile[threadIdx.x][i];
  }
}

}  // namespace ls

using namespace ls;

// scratch for the partial sums + tickets, owned by the library (one per device, grown on demand; single stream use)
namespace ls {
struct GnScratch {
  ScratchBlock partial_b, tickets_b;
  float* partial() const { return reinterpret_cast<float*>(partial_b.ptr); }
  unsigned int* tickets() const { return reinterpret_cast<unsigned int*>(tickets_b.ptr); }
};
static GnScratch g_gn[16];
// 4 MB of partial sums / 32 K tickets cover every UNet / VAE shape of this package; larger requests grow (never free)
static int gn_reserve(GnScratch& sc, size_t partial_floats, size_t ntickets, cudaStream_t stream, const char* who) {
  int rc = scratch_reserve(sc.partial_b, partial_floats * sizeof(float), (size_t)4 << 20, false, stream, who);
  if (rc != 0) return rc;
  return scratch_reserve(sc.tickets_b, ntickets * sizeof(unsigned int), (size_t)32768 * sizeof(unsigned int), true, stream,
                         who);
}
}  // namespace ls

extern "C" int ls_groupnorm_stats(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows,
                                  int32_t rows_per_inst, int32_t groups, float* stats, void* stream) {
  const int C = c1 + (x2 ? c2 : 0);
  if (!x2) c2 = 0;
  LS_CHECK(x1 && stats && rows > 0 && rows_per_inst > 0 && rows % rows_per_inst == 0, "ls_groupnorm_stats: bad args");
  LS_CHECK(groups > 0 && groups <= 32 && C % groups == 0 && C % 8 == 0 && c1 % 8 == 0,
           "ls_groupnorm_stats: C=%d groups=%d unsupported", C, groups);
  int ninst, chunks, rpc;
  gn_chunking(rows, rows_per_inst, 444, ninst, chunks, rpc);
  int dev = 0;
  LS_CUDA(cudaGetDevice(&dev));
  LS_CHECK(dev >= 0 && dev < 16, "ls_groupnorm_stats: device index %d out of range", dev);
  GnScratch& sc = g_gn[dev];
  const size_t need = (size_t)ninst * chunks * groups * 2;
  {
    const int rc = gn_reserve(sc, need, (size_t)ninst, (cudaStream_t)stream, "ls_groupnorm_stats");
    if (rc != 0) return rc;
  }
  const int threads = 256;
  const int nvec = C / 8;
  const int RL = nvec <= threads ? threads / nvec : 1;
  const size_t smem = (size_t)2 * RL * C * sizeof(float);
  LS_CHECK(smem <= 48 * 1024, "ls_groupnorm_stats: C=%d needs %zu bytes of smem", C, smem);
  LS_CUDA(launch_k(gn_stats_kernel, dim3(dim3(chunks, ninst)), dim3(threads), (size_t)(smem), (cudaStream_t)((cudaStream_t)stream), 
      (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, rpc, groups, sc.partial(), stats, sc.tickets()));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_groupnorm_apply(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows,
                                  int32_t rows_per_inst, int32_t groups, const float* stats, const float* gamma,
                                  const float* beta, float eps, int32_t silu, void* y, void* stream) {
  if (!x2) c2 = 0;
  const int C = c1 + c2;
  LS_CHECK(x1 && stats && gamma && beta && y && rows > 0 && rows_per_inst > 0 && rows % rows_per_inst == 0,
           "ls_groupnorm_apply: bad args");
  LS_CHECK(groups > 0 && C % groups == 0 && C % 8 == 0 && c1 % 8 == 0, "ls_groupnorm_apply: C=%d unsupported", C);
  LS_CHECK(2 * C * sizeof(float) <= 48 * 1024, "ls_groupnorm_apply: C=%d too large", C);
  int ninst, chunks, rpc;
  gn_chunking(rows, rows_per_inst, 2368, ninst, chunks, rpc);
  LS_CUDA(launch_k(gn_apply_kernel, dim3(dim3(chunks, ninst)), dim3(256), (size_t)(2 * C * sizeof(float)), (cudaStream_t)((cudaStream_t)stream), 
      (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, rpc, groups, stats, gamma, beta, eps, silu,
      (__half*)y));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}


namespace ls {
// Cluster path of ls_groupnorm: returns 0 = launched, -1 = not applicable (the caller uses the global-rendezvous kernel)
static int groupnorm_cluster_try(const void* x1, int c1, const void* x2, int c2, int64_t rows, int rows_per_inst,
                                 int groups, const float* gamma, const float* beta, float eps, int silu, void* y,
                                 cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("LS_GN_CLUSTER");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return -1;
  const int C = c1 + c2;
  if (C > 5 * GNC_THREADS || groups * 2 > 64) return -1;
  if (c2 != 0 && (c2 % 8) != 0) return -1;
  const int64_t ninst = rows / rows_per_inst;
  if (ninst > (1 << 20)) return -1;
  const int nvec = C / 8;
  const int RL = nvec <= GNC_THREADS ? GNC_THREADS / nvec : 1;
  const size_t red_bytes = (size_t)2 * RL * C * sizeof(float);
  const size_t inst_bytes = (size_t)rows_per_inst * C * 2;
  // smallest cluster whose chunk fits ~96 KB (two CTAs per SM), else ~200 KB (one per SM); then widen while the grid
  // is smaller than the GPU and chunks stay >= 8 rows
  // Measured (tools/gn_bench.py): the cluster kernel wins when many small chunks cover the GPU (per-frame norms: 15.7
  // vs 17.8 us at 32 x 1024 x 320, 8.2 vs 9.5 us at 32 x 64 x 1280) and loses when few CTAs stream > 100 KB each
  // (8x8 joint norm on 16-CTA clusters: 24 vs 15 us), so: chunk <= 96 KB, cluster <= 8, at least 128 CTAs.
  int cl = 0;
  for (int c = 1; c <= 8 && !cl; c *= 2)
    if (rows_per_inst % c == 0 && inst_bytes / c + red_bytes <= (size_t)96 * 1024) cl = c;
  if (!cl) return -1;
  while (cl < 8 && ninst * cl < 128 && (rows_per_inst % (cl * 2)) == 0 && rows_per_inst / (cl * 2) >= 8) cl *= 2;
  if (ninst * cl < 128 && ninst * cl < 64) return -1;
  const int rows_per_cta = rows_per_inst / cl;
  const size_t smem = (size_t)rows_per_cta * C * 2 + red_bytes;
  int dev_i = 0;
  cudaGetDevice(&dev_i);
  if (dev_i < 0 || dev_i >= 16) return -1;
  static bool attr_set_dev[16] = {};  // function attributes are per device
  bool& attr_set = attr_set_dev[dev_i];
  static bool broken[5] = {false, false, false, false, false};  // per log2(cl): a launch failed once -> never retry
  int lg = 0;
  while ((1 << lg) < cl) ++lg;
  if (broken[lg]) return -1;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      enabled = 0;
      return -1;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ninst * cl));
  cfg.blockDim = dim3(GNC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gn_cluster_kernel, (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst,
                                     rows_per_cta, cl, groups, gamma, beta, eps, silu, (__half*)y);
  if (e != cudaSuccess) {
    cudaGetLastError();
    broken[lg] = true;
    return -1;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
}  // namespace ls

// Fused single-launch GroupNorm (+SiLU): y = GN(x) [* sigmoid]; replaces the stats + apply pair when the grid fits the
// GPU in one co-resident wave (always true for the UNet / VAE shapes); otherwise falls back to the two-kernel path.
extern "C" int ls_groupnorm(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows, int32_t rows_per_inst,
                            int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                            float* stats_scratch, void* y, void* stream) {
  if (!x2) c2 = 0;
  const int C = c1 + c2;
  LS_CHECK(x1 && gamma && beta && y && stats_scratch && rows > 0 && rows_per_inst > 0 && rows % rows_per_inst == 0,
           "ls_groupnorm: bad args");
  LS_CHECK(groups > 0 && groups <= 32 && C % groups == 0 && C % 8 == 0 && c1 % 8 == 0,
           "ls_groupnorm: C=%d groups=%d unsupported", C, groups);
  {
    // small instances: one CTA per (instance, group), no cross-CTA exchange (gn_group_kernel).  LS_GN_GROUP=0 disables.
    static int env_grp = -1;
    if (env_grp < 0) {
      const char* e = getenv("LS_GN_GROUP");
      env_grp = e ? atoi(e) : 1;
    }
    const int cg = C / groups;
    const int64_t ninst_g = rows / rows_per_inst;
    if (env_grp && cg % 4 == 0 && c1 % 4 == 0 && (int64_t)rows_per_inst * cg * 2 <= 48 * 1024 && ninst_g * groups >= 32 &&
        ninst_g <= 65535 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0) {
      LS_CUDA(launch_k(gn_group_kernel, dim3(groups, (unsigned)ninst_g), dim3(256), (size_t)0, (cudaStream_t)stream,
                       (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, groups, gamma, beta, eps, silu,
                       (__half*)y));
      LS_CUDA(cudaGetLastError());
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      return 0;
    }
  }
  if (ls::groupnorm_cluster_try(x1, c1, x2, c2, rows, rows_per_inst, groups, gamma, beta, eps, silu, y,
                                (cudaStream_t)stream) == 0)
    return 0;
  const int threads = 256;
  const int nvec = C / 8;
  const int RL = nvec <= threads ? threads / nvec : 1;
  size_t smem = (size_t)2 * RL * C * sizeof(float);
  if (smem < (size_t)2 * C * sizeof(float)) smem = (size_t)2 * C * sizeof(float);
  int dev = 0;
  LS_CUDA(cudaGetDevice(&dev));
  static int occ_cache[16][8] = {};  // per device, per smem bucket (8 KB steps)
  int sms = 0;
  LS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int bucket = (int)((smem + 8191) / 8192);
  int occ = 0;
  if (dev < 16 && bucket < 8) {
    if (occ_cache[dev][bucket] == 0) {
      int o = 0;
      LS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, gn_fused_kernel, threads, (size_t)bucket * 8192));
      occ_cache[dev][bucket] = o > 0 ? o : -1;
    }
    occ = occ_cache[dev][bucket];
  }
  const int ninst0 = (int)(rows / rows_per_inst);
  const int capacity = occ > 0 ? occ * sms : 0;
  if (smem > 48 * 1024 || capacity < ninst0 || dev >= 16 || C > 2560) {
    // cannot guarantee co-residency: two launches
    int rc = ls_groupnorm_stats(x1, c1, x2, c2, rows, rows_per_inst, groups, stats_scratch, stream);
    if (rc != 0) return rc;
    return ls_groupnorm_apply(x1, c1, x2, c2, rows, rows_per_inst, groups, stats_scratch, gamma, beta, eps, silu, y,
                              stream);
  }
  int ninst, chunks, rpc;
  int target = capacity < 592 ? capacity : 592;
  gn_chunking(rows, rows_per_inst, target, ninst, chunks, rpc);
  while (chunks > 1 && (int64_t)chunks * ninst > capacity) {  // gn_chunking rounds up: stay inside one wave
    rpc += 1;
    chunks = (rows_per_inst + rpc - 1) / rpc;
  }
  LS_CHECK((int64_t)chunks * ninst <= capacity, "ls_groupnorm: %d instances do not fit one wave", ninst);
  GnScratch& sc = g_gn[dev];
  const size_t need = (size_t)ninst * chunks * groups * 2;
  const size_t ntick = (size_t)2 * ninst + 2;
  {
    const int rc = gn_reserve(sc, need, ntick, (cudaStream_t)stream, "ls_groupnorm");
    if (rc != 0) return rc;
  }
  // cooperative launch: the ticket rendezvous inside gn_fused_kernel needs the whole grid resident (chunks * ninst <=
  // capacity is checked above against the occupancy query; the driver now enforces it against whatever else runs)
  LS_CUDA(launch_coop_k(gn_fused_kernel, dim3(chunks, ninst), dim3(threads), smem, (cudaStream_t)stream,
                   (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, rpc, groups, sc.partial(), sc.tickets(),
                   gamma, beta, eps, silu, (__half*)y));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// GroupNorm (+ SiLU) of the virtual concatenation [x1 | x2] from the partials their producing GEMMs wrote
// (LsGemmArgs.gn_partials_out): no statistics pass, no rendezvous, one launch.
extern "C" int ls_groupnorm_parts(const void* x1, int32_t c1, const float* parts1, int32_t ld1, const void* x2, int32_t c2,
                                  const float* parts2, int32_t ld2, int64_t rows, int32_t rows_per_inst, int32_t groups,
                                  int32_t unit, const float* gamma, const float* beta, float eps, int32_t silu, void* y,
                                  void* stream) {
  if (!x2) c2 = 0;
  const int C = c1 + c2;
  LS_CHECK(x1 && parts1 && gamma && beta && y && rows > 0 && rows_per_inst > 0 && rows % rows_per_inst == 0 &&
               (c2 == 0 || parts2 != nullptr),
           "ls_groupnorm_parts: bad args");
  LS_CHECK(groups > 0 && groups <= 32 && C % groups == 0 && C % 8 == 0 && c1 % 8 == 0,
           "ls_groupnorm_parts: C=%d groups=%d unsupported", C, groups);
  const int cg = C / groups;
  LS_CHECK(rows_per_inst % 128 == 0 && unit >= 1 && cg % unit == 0 && c1 % unit == 0 && ld1 >= c1 / unit &&
               (c2 == 0 || ld2 >= c2 / unit),
           "ls_groupnorm_parts: rows_per_inst %d must be a multiple of 128, unit %d must divide C / groups = %d and c1 = %d",
           rows_per_inst, unit, cg, c1);
  LS_CHECK((reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
           "ls_groupnorm_parts: gamma / beta must be 16-byte aligned");
  // channel split: whole groups and whole 8-channel vectors per CTA, at least 240 channels each
  int csplit = 1;
  for (int s2 = 8; s2 > 1; s2 >>= 1) {
    if (C % s2 != 0) continue;
    const int cl = C / s2;
    if (cl >= 240 && cl % cg == 0 && cl % 8 == 0) {
      csplit = s2;
      break;
    }
  }
  // one wave of three CTAs per SM: with more, the second round pays the per-CTA latency chain (partials -> scale / shift
  // -> first loads) again
  int ninst, chunks, rpc;
  int dev = 0, sms = 148;
  LS_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int target = 3 * sms / csplit;
  gn_chunking(rows, rows_per_inst, target < 1 ? 1 : target, ninst, chunks, rpc);
  while (chunks > 1 && (int64_t)chunks * csplit * ninst > 3 * sms) {  // gn_chunking rounds up: stay inside the wave
    rpc += 1;
    chunks = (rows_per_inst + rpc - 1) / rpc;
  }
  LS_CUDA(launch_k(gn_parts_kernel, dim3(chunks * csplit, ninst), dim3(256), (size_t)0, (cudaStream_t)stream,
                   (const __half*)x1, c1, (const float2*)parts1, ld1, (const __half*)x2, c2, (const float2*)parts2, ld2,
                   rows_per_inst, rpc, groups, unit, csplit, gamma, beta, eps, silu, (__half*)y));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <int VPL, int ROWS>
static int launch_ln(const void* x, int64_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                      const float* pe, int32_t rows_per_frame, int32_t nframes, void* y, cudaStream_t stream) {
  const int wpb = 8;
  const int64_t rows_per_block = (int64_t)wpb * ROWS;
  LS_CUDA(launch_k(layernorm_kernel<VPL, ROWS>, dim3((unsigned)((rows + rows_per_block - 1) / rows_per_block)), dim3(wpb * 32), (size_t)(0), (cudaStream_t)(stream), 
      (const __half*)x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, (__half*)y));
  return 0;
}

extern "C" int ls_layernorm(const void* x, int64_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                            const float* pe, int32_t rows_per_frame, int32_t nframes, void* y, void* stream) {
  LS_CHECK(x && y && gamma && beta && rows > 0, "ls_layernorm: bad args");
  LS_CHECK(C % 8 == 0 && C <= 5 * 256, "ls_layernorm: C=%d unsupported (multiple of 8, <= 1280)", C);
  LS_CHECK(pe == nullptr || (rows_per_frame > 0 && nframes > 0), "ls_layernorm: bad pe geometry");
  LS_CHECK(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
             reinterpret_cast<uintptr_t>(pe)) & 15) == 0, "ls_layernorm: gamma/beta/pe must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = C / 8;
  int rc;
  static int env_ln40 = -1;
  if (env_ln40 < 0) {
    const char* e = getenv("LS_LN40");  // 0: always the generic kernel (A/B measurements)
    env_ln40 = e ? atoi(e) : 1;
  }
  const int lpr = (C % 40 == 0) ? C / 40 : 0;
  if (env_ln40 && (lpr == 8 || lpr == 16 || lpr == 32)) {
    // two row groups per warp (10 loads in flight per lane) only when that still leaves >= 2 blocks per SM
    const int64_t rpb2 = 8 * (32 / lpr) * 2;
    const bool two = (rows + rpb2 - 1) / rpb2 >= 2 * 148;
    const int64_t rpb = two ? rpb2 : rpb2 / 2;
    const dim3 grid((unsigned)((rows + rpb - 1) / rpb)), block(256);
#define LS_LN40(LPR_, P_)                                                                                          \
  LS_CUDA(launch_k(layernorm40_kernel<LPR_, P_>, grid, block, (size_t)0, st, (const __half*)x, rows, C, gamma, beta, \
                   eps, pe, rows_per_frame, nframes, (__half*)y))
    if (lpr == 8) {
      if (two) LS_LN40(8, 2); else LS_LN40(8, 1);
    } else if (lpr == 16) {
      if (two) LS_LN40(16, 2); else LS_LN40(16, 1);
    } else {
      if (two) LS_LN40(32, 2); else LS_LN40(32, 1);
    }
#undef LS_LN40
    LS_CUDA(cudaGetLastError());
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  if (nvec <= 32) rc = launch_ln<1, 8>(x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, y, st);
  else if (nvec <= 64) rc = launch_ln<2, 4>(x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, y, st);
  else if (nvec <= 96) rc = launch_ln<3, 2>(x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, y, st);
  else rc = launch_ln<5, 1>(x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, y, st);
  if (rc != 0) return rc;
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_softmax_rows(const float* s, int64_t rows, int32_t cols, float scale, void* p, void* stream) {
  LS_CHECK(s && p && rows > 0 && cols % 8 == 0 && cols <= SM_MAXV * 256 && scale > 0.f,
           "ls_softmax_rows: cols=%d unsupported", cols);
  const int wpb = 8;
  LS_CUDA(launch_k(softmax_rows_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), 
      s, rows, cols, scale, (__half*)p));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_transpose(const void* x, int32_t batch, int32_t R, int32_t C, void* y, void* stream) {
  LS_CHECK(x && y && batch > 0 && R > 0 && C > 0, "ls_transpose: bad args");
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  LS_CUDA(launch_k(transpose_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), (const __half*)x, R, C, (__half*)y));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
