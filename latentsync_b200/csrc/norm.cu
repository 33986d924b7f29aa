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
// Deterministic two-level reduction (no floating-point atomics): every CTA reduces its row chunk per channel pair,
// folds pairs into groups in a fixed order and writes one partial per (instance, chunk, group); the CTA that arrives
// last at the per-instance ticket (an integer atomic) sums the partials in chunk order and resets the ticket.
__global__ void gn_stats_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2,
                                int rows_per_inst, int rows_per_chunk, int groups, float* __restrict__ partial,
                                float* __restrict__ stats, unsigned int* __restrict__ tickets) {
  extern __shared__ float gs_sm[];  // [npairs] sums, [npairs] sums of squares
  __shared__ int s_last;
  const int C = c1 + c2;
  const int cg = C / groups;
  const int npairs = C >> 1;
  float* sh_s = gs_sm;
  float* sh_q = gs_sm + npairs;
  const int inst = blockIdx.y;
  const int chunks = gridDim.x;
  const int64_t row0 = (int64_t)inst * rows_per_inst + (int64_t)blockIdx.x * rows_per_chunk;
  int64_t row_end = row0 + rows_per_chunk;
  const int64_t inst_end = (int64_t)(inst + 1) * rows_per_inst;
  if (row_end > inst_end) row_end = inst_end;
  for (int pair = threadIdx.x; pair < npairs; pair += blockDim.x) {
    const int c = pair * 2;
    const __half* src;
    int ld, cc;
    if (c < c1) {
      src = x1;
      ld = c1;
      cc = c;
    } else {
      src = x2;
      ld = c2;
      cc = c - c1;
    }
    float s = 0.f, ss = 0.f;
#pragma unroll 4
    for (int64_t row = row0; row < row_end; ++row) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(src + row * ld + cc));
      s += f.x + f.y;
      ss += f.x * f.x + f.y * f.y;
    }
    sh_s[pair] = s;
    sh_q[pair] = ss;
  }
  __syncthreads();
  float* my_partial = partial + ((int64_t)inst * chunks + blockIdx.x) * groups * 2;
  if (threadIdx.x < groups) {
    const int ppg = cg >> 1;
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < ppg; ++i) {
      s += sh_s[threadIdx.x * ppg + i];
      ss += sh_q[threadIdx.x * ppg + i];
    }
    my_partial[threadIdx.x * 2] = s;
    my_partial[threadIdx.x * 2 + 1] = ss;
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

__global__ void gn_apply_kernel(const __half* __restrict__ x1, int c1, const __half* __restrict__ x2, int c2,
                                int rows_per_inst, int rows_per_chunk, int groups, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                __half* __restrict__ y) {
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
    const uint4 u = *reinterpret_cast<const uint4*>(src);
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
// one warp per row, values held in registers, exact two-pass variance
constexpr int LN_MAXV = 5;  // C <= 5 * 32 * 8 = 1280

__global__ void layernorm_kernel(const __half* __restrict__ x, int64_t rows, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, const float* __restrict__ pe,
                                 int rows_per_frame, int nframes, __half* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 3;
  float v[LN_MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + vi * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        v[i][2 * e] = f.x;
        v[i][2 * e + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  const float* pe_row = nullptr;
  if (pe != nullptr) pe_row = pe + (int64_t)((row / rows_per_frame) % nframes) * C;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < nvec) {
      const int c = vi * 8;
      float r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        r[e] = (v[i][e] - mean) * rstd * __ldg(gamma + c + e) + __ldg(beta + c + e);
        if (pe_row != nullptr) r[e] += __ldg(pe_row + c + e);
      }
      uint4 w;
      __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(r[2 * e], r[2 * e + 1]);
      *reinterpret_cast<uint4*>(y + row * C + c) = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------- row softmax
constexpr int SM_MAXV = 8;  // cols <= 8 * 32 * 8 = 2048

__global__ void softmax_rows_kernel(const float* __restrict__ s, int64_t rows, int cols, float scale,
                                    __half* __restrict__ p) {
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
        v[i][e] = exp2f((v[i][e] - mx) * sl);
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
  float* partial = nullptr;
  size_t partial_floats = 0;
  unsigned int* tickets = nullptr;
  size_t ntickets = 0;
};
static GnScratch g_gn[16];
}  // namespace ls

extern "C" int ls_groupnorm_stats(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows,
                                  int32_t rows_per_inst, int32_t groups, float* stats, void* stream) {
  const int C = c1 + (x2 ? c2 : 0);
  if (!x2) c2 = 0;
  LS_CHECK(x1 && stats && rows > 0 && rows_per_inst > 0 && rows % rows_per_inst == 0, "ls_groupnorm_stats: bad args");
  LS_CHECK(groups > 0 && groups <= 32 && C % groups == 0 && (C / groups) % 2 == 0 && c1 % 2 == 0,
           "ls_groupnorm_stats: C=%d groups=%d unsupported", C, groups);
  int ninst, chunks, rpc;
  gn_chunking(rows, rows_per_inst, 1184, ninst, chunks, rpc);
  int dev = 0;
  LS_CUDA(cudaGetDevice(&dev));
  LS_CHECK(dev >= 0 && dev < 16, "ls_groupnorm_stats: device index %d out of range", dev);
  GnScratch& sc = g_gn[dev];
  const size_t need = (size_t)ninst * chunks * groups * 2;
  if (need > sc.partial_floats || (size_t)ninst > sc.ntickets) {
    // growing is not stream-ordered: only legal outside graph capture (plans warm up eagerly before capturing)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing((cudaStream_t)stream, &cs);
    LS_CHECK(cs == cudaStreamCaptureStatusNone, "ls_groupnorm_stats: scratch must be sized by an eager warm-up run");
    LS_CUDA(cudaDeviceSynchronize());
    if (need > sc.partial_floats) {
      if (sc.partial) cudaFree(sc.partial);
      sc.partial_floats = need > (1u << 20) ? need : (1u << 20);
      LS_CUDA(cudaMalloc(&sc.partial, sc.partial_floats * sizeof(float)));
    }
    if ((size_t)ninst > sc.ntickets) {
      if (sc.tickets) cudaFree(sc.tickets);
      sc.ntickets = ninst > 4096 ? ninst : 4096;
      LS_CUDA(cudaMalloc(&sc.tickets, sc.ntickets * sizeof(unsigned int)));
      LS_CUDA(cudaMemset(sc.tickets, 0, sc.ntickets * sizeof(unsigned int)));
    }
  }
  int threads = C / 2;
  if (threads > 512) threads = 512;
  threads = (threads + 31) / 32 * 32;
  if (threads < 256) threads = 256;  // the final reduction uses 4 threads per (group, stat): 4 * 64 = 256
  const size_t smem = (size_t)C * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, ninst), threads, smem, (cudaStream_t)stream>>>(
      (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, rpc, groups, sc.partial, stats, sc.tickets);
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
  gn_apply_kernel<<<dim3(chunks, ninst), 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const __half*)x1, c1, (const __half*)x2, c2, rows_per_inst, rpc, groups, stats, gamma, beta, eps, silu,
      (__half*)y);
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_layernorm(const void* x, int64_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                            const float* pe, int32_t rows_per_frame, int32_t nframes, void* y, void* stream) {
  LS_CHECK(x && y && gamma && beta && rows > 0, "ls_layernorm: bad args");
  LS_CHECK(C % 8 == 0 && C <= LN_MAXV * 256, "ls_layernorm: C=%d unsupported (multiple of 8, <= 1280)", C);
  LS_CHECK(pe == nullptr || (rows_per_frame > 0 && nframes > 0), "ls_layernorm: bad pe geometry");
  const int wpb = 8;
  layernorm_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      (const __half*)x, rows, C, gamma, beta, eps, pe, rows_per_frame, nframes, (__half*)y);
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_softmax_rows(const float* s, int64_t rows, int32_t cols, float scale, void* p, void* stream) {
  LS_CHECK(s && p && rows > 0 && cols % 8 == 0 && cols <= SM_MAXV * 256 && scale > 0.f,
           "ls_softmax_rows: cols=%d unsupported", cols);
  const int wpb = 8;
  softmax_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      s, rows, cols, scale, (__half*)p);
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int ls_transpose(const void* x, int32_t batch, int32_t R, int32_t C, void* y, void* stream) {
  LS_CHECK(x && y && batch > 0 && R > 0 && C > 0, "ls_transpose: bad args");
  dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const __half*)x, R, C, (__half*)y);
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}
