// Fused flash-style attention for the UNet (spatial self-attention, audio cross-attention, temporal attention).
// fp16 operands on the warp-level tensor-core path (mma.sync m16n8k16), fp32 online softmax and accumulation.
// Both CFG halves and all heads run in one launch.  Temporal attention gathers its 16 frames straight from the
// (b f) x HW token matrix with a row stride (no "(b f) s c -> (b s) f c" permute copies).
//
// Replaces F.scaled_dot_product_attention at latentsync/models/attention.py:271 and motion_module.py:300
// (and the split_heads/concat_heads permutes at attention.py:238-248).
// NOTE: SDPA is 3.4 % of the UNet FLOPs; the 96 % GEMM/conv share runs on tcgen05 (gemm_tc.cu).
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>

namespace ls {

extern std::atomic<int64_t> g_launch_count;

struct AttnKParams {
  const __half* q;
  const __half* k;
  const __half* v;
  __half* o;
  int ldq, ldk, ldv, ldo;
  int heads, sq, skv, batch;
  int q_inner;
  int64_t q_outer, q_in_stride, q_seq;
  int kv_inner;
  int64_t kv_outer, kv_in_stride, kv_seq;
  float scale_log2;  // softmax scale * log2(e), applied to the fp32 scores
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr float kLog2e = 1.4426950408889634f;

// -------------------------------------------------------------------------------------------------------------
// General kernel: 64 query rows per CTA (4 warps x 16 rows), 64-key tiles double buffered with cp.async.
// -------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnKParams p) {
  pdl_prologue();
  constexpr int DP = (D + 15) / 16 * 16;
  constexpr int DS = DP + 8;  // padded smem row (halfs): conflict-free ldmatrix, 16B aligned
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  constexpr int KS = DP / 16;
  constexpr int NT_O = DP / 8;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __half* sQ = reinterpret_cast<__half*>(attn_smem);
  __half* sK = sQ + 64 * DS;
  __half* sV = sK + 2 * 64 * DS;

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float sl = p.scale_log2;
  const int64_t qbase = (int64_t)(b / p.q_inner) * p.q_outer + (int64_t)(b % p.q_inner) * p.q_in_stride;
  const int64_t kvbase = (int64_t)(b / p.kv_inner) * p.kv_outer + (int64_t)(b % p.kv_inner) * p.kv_in_stride;

  if constexpr (DP > D) {
    // zero the pad columns once; cp.async never writes them
    for (int i = tid; i < 5 * 64; i += 128) {
      __half* row = sQ + i * DS;
      for (int c = D; c < DP; ++c) row[c] = __float2half(0.f);
    }
  }
  for (int i = tid; i < 64 * CPR; i += 128) {
    const int r = i / CPR, c = i - r * CPR;
    const int qrow = qt * 64 + r;
    const bool ok = qrow < p.sq;
    const __half* src = p.q + (qbase + (int64_t)(ok ? qrow : 0) * p.q_seq) * p.ldq + h * D + c * 8;
    cp_async16(sQ + r * DS + c * 8, src, ok);
  }
  auto load_kv = [&](int kt, int buf) {
    for (int i = tid; i < 64 * CPR; i += 128) {
      const int r = i / CPR, c = i - r * CPR;
      const int krow = kt * 64 + r;
      const bool ok = krow < p.skv;
      const int64_t row = kvbase + (int64_t)(ok ? krow : 0) * p.kv_seq;
      cp_async16(sK + (buf * 64 + r) * DS + c * 8, p.k + row * p.ldk + h * D + c * 8, ok);
      cp_async16(sV + (buf * 64 + r) * DS + c * 8, p.v + row * p.ldv + h * D + c * 8, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[KS][4];
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  const int nkt = (p.skv + 63) / 64;
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_kv(kt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(qf[ks], sQ + row * DS + col);
      }
    }
    const __half* bK = sK + buf * 64 * DS;
    const __half* bV = sV + buf * 64 * DS;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(bf, bK + key * DS + col);
        mma_16816(s[2 * np], qf[ks], bf[0], bf[1]);
        mma_16816(s[2 * np + 1], qf[ks], bf[2], bf[3]);
      }
    }
    // mask keys beyond skv (only possible in the last tile)
    if (kt * 64 + 64 > p.skv) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kt * 64 + nt * 8 + 2 * t;
        if (key >= p.skv) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= p.skv) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float al_lo = mufu_ex2((m_lo - mn_lo) * sl), al_hi = mufu_ex2((m_hi - mn_hi) * sl);
    m_lo = mn_lo;
    m_hi = mn_hi;
    float rs_lo = 0.f, rs_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = mufu_ex2((s[nt][0] - mn_lo) * sl);
      s[nt][1] = mufu_ex2((s[nt][1] - mn_lo) * sl);
      s[nt][2] = mufu_ex2((s[nt][2] - mn_hi) * sl);
      s[nt][3] = mufu_ex2((s[nt][3] - mn_hi) * sl);
      rs_lo += s[nt][0] + s[nt][1];
      rs_hi += s[nt][2] + s[nt][3];
    }
    l_lo = l_lo * al_lo + rs_lo;
    l_hi = l_hi * al_hi + rs_hi;
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o[i][0] *= al_lo;
      o[i][1] *= al_lo;
      o[i][2] *= al_hi;
      o[i][3] *= al_hi;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < KS; ++dp) {
        uint32_t bf[4];
        const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = dp * 16 + (lane >> 4) * 8;
        ldsm_x4_trans(bf, bV + key * DS + col);
        mma_16816(o[2 * dp], pa, bf[0], bf[1]);
        mma_16816(o[2 * dp + 1], pa, bf[2], bf[3]);
      }
    }
    __syncthreads();
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
  const int row_lo = qt * 64 + warp * 16 + g, row_hi = row_lo + 8;
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    const int col = nt * 8 + 2 * t;
    if (col < D) {
      if (row_lo < p.sq) {
        __half* dst = p.o + (qbase + (int64_t)row_lo * p.q_seq) * p.ldo + h * D + col;
        *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
      }
      if (row_hi < p.sq) {
        __half* dst = p.o + (qbase + (int64_t)row_hi * p.q_seq) * p.ldo + h * D + col;
        *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// Short-sequence kernel (sq, skv <= 16): one warp per (batch, head) problem, 4 problems per CTA.
// Used for temporal attention over the 16 frames and for the 4x4 mid-block spatial attention.
// -------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_short_kernel(const AttnKParams p) {
  pdl_prologue();
  constexpr int DP = (D + 15) / 16 * 16;
  constexpr int DS = DP + 8;
  constexpr int CPR = D / 8;
  constexpr int KS = DP / 16;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float sl = p.scale_log2;
  __half* sQ = reinterpret_cast<__half*>(attn_smem) + warp * 3 * 16 * DS;
  __half* sK = sQ + 16 * DS;
  __half* sV = sK + 16 * DS;
  const int64_t prob = (int64_t)blockIdx.x * 4 + warp;  // problem = b * heads + h
  if (prob >= (int64_t)p.batch * p.heads) return;
  const int b = (int)(prob / p.heads), h = (int)(prob % p.heads);
  const int64_t qbase = (int64_t)(b / p.q_inner) * p.q_outer + (int64_t)(b % p.q_inner) * p.q_in_stride;
  const int64_t kvbase = (int64_t)(b / p.kv_inner) * p.kv_outer + (int64_t)(b % p.kv_inner) * p.kv_in_stride;

  if constexpr (DP > D) {
    for (int i = lane; i < 48; i += 32) {
      __half* row = sQ + i * DS;
      for (int c = D; c < DP; ++c) row[c] = __float2half(0.f);
    }
  }
  for (int i = lane; i < 16 * CPR; i += 32) {
    const int r = i / CPR, c = i - r * CPR;
    const bool okq = r < p.sq, okk = r < p.skv;
    const int64_t qrow = qbase + (int64_t)(okq ? r : 0) * p.q_seq;
    const int64_t krow = kvbase + (int64_t)(okk ? r : 0) * p.kv_seq;
    cp_async16(sQ + r * DS + c * 8, p.q + qrow * p.ldq + h * D + c * 8, okq);
    cp_async16(sK + r * DS + c * 8, p.k + krow * p.ldk + h * D + c * 8, okk);
    cp_async16(sV + r * DS + c * 8, p.v + krow * p.ldv + h * D + c * 8, okk);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4], bf[4];
    ldsm_x4(a, sQ + ((lane & 7) + ((lane >> 3) & 1) * 8) * DS + ks * 16 + (lane >> 4) * 8);
    ldsm_x4(bf, sK + ((lane & 7) + (lane >> 4) * 8) * DS + ks * 16 + ((lane >> 3) & 1) * 8);
    mma_16816(s[0], a, bf[0], bf[1]);
    mma_16816(s[1], a, bf[2], bf[3]);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int key = nt * 8 + 2 * t;
    if (key >= p.skv) s[nt][0] = s[nt][2] = -INFINITY;
    if (key + 1 >= p.skv) s[nt][1] = s[nt][3] = -INFINITY;
  }
  float mx_lo = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
  float mx_hi = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s[nt][0] = mufu_ex2((s[nt][0] - mx_lo) * sl);
    s[nt][1] = mufu_ex2((s[nt][1] - mx_lo) * sl);
    s[nt][2] = mufu_ex2((s[nt][2] - mx_hi) * sl);
    s[nt][3] = mufu_ex2((s[nt][3] - mx_hi) * sl);
    l_lo += s[nt][0] + s[nt][1];
    l_hi += s[nt][2] + s[nt][3];
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
  uint32_t pa[4];
  pa[0] = pack_h2(s[0][0] * inv_lo, s[0][1] * inv_lo);
  pa[1] = pack_h2(s[0][2] * inv_hi, s[0][3] * inv_hi);
  pa[2] = pack_h2(s[1][0] * inv_lo, s[1][1] * inv_lo);
  pa[3] = pack_h2(s[1][2] * inv_hi, s[1][3] * inv_hi);
#pragma unroll
  for (int dp = 0; dp < KS; ++dp) {
    uint32_t bf[4];
    ldsm_x4_trans(bf, sV + ((lane & 7) + ((lane >> 3) & 1) * 8) * DS + dp * 16 + (lane >> 4) * 8);
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    mma_16816(o0, pa, bf[0], bf[1]);
    mma_16816(o1, pa, bf[2], bf[3]);
#pragma unroll
    for (int half_ = 0; half_ < 2; ++half_) {
      const float* oo = half_ ? o1 : o0;
      const int col = dp * 16 + half_ * 8 + 2 * t;
      if (col < D) {
        if (g < p.sq) {
          __half* dst = p.o + (qbase + (int64_t)g * p.q_seq) * p.ldo + h * D + col;
          *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(oo[0], oo[1]);
        }
        if (g + 8 < p.sq) {
          __half* dst = p.o + (qbase + (int64_t)(g + 8) * p.q_seq) * p.ldo + h * D + col;
          *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(oo[2], oo[3]);
        }
      }
    }
  }
}

template <int D>
static int launch_attn(const AttnKParams& p, cudaStream_t stream) {
  constexpr int DP = (D + 15) / 16 * 16;
  constexpr int DS = DP + 8;
  if (p.sq <= 16 && p.skv <= 16) {
    const size_t smem = (size_t)4 * 3 * 16 * DS * sizeof(__half);
    static bool set_short_dev[16] = {};
    bool& set_short = set_short_dev[dev_slot()];
    if (!set_short) {
      LS_CUDA(cudaFuncSetAttribute(attn_short_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_short = true;
    }
    const int64_t probs = (int64_t)p.batch * p.heads;
    LS_CUDA(launch_k(attn_short_kernel<D>, dim3((unsigned)((probs + 3) / 4)), dim3(128), (size_t)(smem), (cudaStream_t)(stream), p));
  } else {
    const size_t smem = (size_t)5 * 64 * DS * sizeof(__half);
    static bool set_gen_dev[16] = {};
    bool& set_gen = set_gen_dev[dev_slot()];
    if (!set_gen) {
      LS_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_gen = true;
    }
    dim3 grid((p.sq + 63) / 64, p.heads, p.batch);
    LS_CUDA(launch_k(attn_fwd_kernel<D>, dim3(grid), dim3(128), (size_t)(smem), (cudaStream_t)(stream), p));
  }
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int attention_tc_try(const LsAttnArgs* a, cudaStream_t stream);  // attention_tc.cu (tcgen05 path, sq >= 128)

static int attention_impl(const LsAttnArgs* a, cudaStream_t stream) {
  LS_CHECK(a && a->q && a->k && a->v && a->out, "ls_attention: null pointer");
  LS_CHECK(a->batch > 0 && a->heads > 0 && a->sq > 0 && a->skv > 0, "ls_attention: bad sizes");
  LS_CHECK(a->batch <= 65535 || (a->sq <= 16 && a->skv <= 16), "ls_attention: batch %d too large for grid.z", a->batch);
  LS_CHECK((a->ldq % 8 == 0) && (a->ldk % 8 == 0) && (a->ldv % 8 == 0) && (a->ldo % 2 == 0),
           "ls_attention: leading dims must be multiples of 8");
  LS_CHECK(a->q_inner >= 1 && a->kv_inner >= 1, "ls_attention: inner must be >= 1");
  LS_CHECK(a->scale > 0.f, "ls_attention: scale must be > 0");
  {
    const int rc = attention_tc_try(a, stream);
    if (rc >= 0) return rc;
  }
  AttnKParams p;
  p.q = reinterpret_cast<const __half*>(a->q);
  p.k = reinterpret_cast<const __half*>(a->k);
  p.v = reinterpret_cast<const __half*>(a->v);
  p.o = reinterpret_cast<__half*>(a->out);
  p.ldq = a->ldq;
  p.ldk = a->ldk;
  p.ldv = a->ldv;
  p.ldo = a->ldo;
  p.heads = a->heads;
  p.sq = a->sq;
  p.skv = a->skv;
  p.batch = a->batch;
  p.q_inner = a->q_inner;
  p.q_outer = a->q_outer_stride;
  p.q_in_stride = a->q_inner_stride;
  p.q_seq = a->q_seq_stride;
  p.kv_inner = a->kv_inner;
  p.kv_outer = a->kv_outer_stride;
  p.kv_in_stride = a->kv_inner_stride;
  p.kv_seq = a->kv_seq_stride;
  p.scale_log2 = a->scale * kLog2e;
  switch (a->head_dim) {
    case 40: return launch_attn<40>(p, stream);
    case 80: return launch_attn<80>(p, stream);
    case 160: return launch_attn<160>(p, stream);
    case 16: return launch_attn<16>(p, stream);
    case 32: return launch_attn<32>(p, stream);
    case 64: return launch_attn<64>(p, stream);
    default: ls::set_error("ls_attention: unsupported head_dim %d", a->head_dim); return 1;
  }
}

}  // namespace ls

extern "C" int ls_attention(const LsAttnArgs* args, void* stream) {
  return ls::attention_impl(args, reinterpret_cast<cudaStream_t>(stream));
}
