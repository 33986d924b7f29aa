// Shared device/host helpers for the latentsync_b200 sm_100a kernels.
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace ls {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define LS_CHECK(cond, ...)                \
  do {                                     \
    if (!(cond)) {                         \
      ls::set_error(__VA_ARGS__);          \
      return 1;                            \
    }                                      \
  } while (0)
#define LS_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ls::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// host: launch helper; LS_PDL=1 in the environment adds the PDL attribute (off by default: measured no gain in graphs)
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Cooperative launch: the driver only starts the grid when ALL its CTAs can be resident at once (and fails the launch
// if they never can).  Used by the kernels whose CTAs wait for one another inside the kernel - the single-launch
// GroupNorm (global ticket rendezvous) and split-K GEMMs - so that their co-residency does not rest on "nothing else is
// running": a concurrent kernel on another stream (NCCL) or an early-launched dependent (PDL) can no longer starve it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// index of the current device, clamped into [0, 16): key of the per-device caches (function attributes such as
// cudaFuncAttributeMaxDynamicSharedMemorySize are per device; a process-wide `static bool` is wrong for cuda:1)
inline int dev_slot() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d > 15) ? 15 : d;
}

// Library-owned device scratch (split-K partials, GroupNorm partial sums, tickets): allocated on first use, grown
// OUTSIDE stream capture only, and a superseded block is never freed - CUDA graphs captured earlier still point at it.
struct ScratchBlock {
  void* ptr = nullptr;
  size_t bytes = 0;
};
// returns 0 on success; `zero`: cudaMemset the new block (tickets)
inline int scratch_reserve(ScratchBlock& b, size_t need, size_t minimum, bool zero, cudaStream_t stream, const char* who) {
  if (need <= b.bytes) return 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cs);
  LS_CHECK(cs == cudaStreamCaptureStatusNone, "%s: scratch must be sized by an eager warm-up run", who);
  LS_CUDA(cudaDeviceSynchronize());
  const size_t bytes = need > minimum ? need : minimum;
  void* p = nullptr;
  LS_CUDA(cudaMalloc(&p, bytes));
  if (zero) LS_CUDA(cudaMemset(p, 0, bytes));
  b.ptr = p;  // the old block (if any) stays allocated for the life of the process: earlier graphs may reference it
  b.bytes = bytes;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (kernel aborts with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("latentsync_b200: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Polite wait for warps that are NOT on the critical path (epilogue warps waiting for an accumulator): back off with
// nanosleep so their polling does not steal issue slots from the TMA-producer / MMA-issuer warps of the same SM
// sub-partition (the warp arbiter favours them otherwise).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(spins < 8 ? 32 : 128);
    if (++spins > (1u << 24)) {
      printf("latentsync_b200: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// TMA tiled loads, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// TMA tiled store smem -> global (bulk async-group completion); out-of-bounds box elements are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16/bf16 operands, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart.
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor; version=1 for sm_100.)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);       // start address, 16-byte units
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}

// true in exactly one lane of the (converged) warp.  Control warps run their loops with all lanes and guard the issue
// instructions with this predicate: under `if (lane == 0)` ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT +
// BRA.U.ANY loop and moves each operand to a uniform register with R2UR.
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile(
      "{\n.reg .pred E;\n"
      "elect.sync _|E, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, E;\n}"
      : "=r"(e));
  return e != 0;
}

// Programmatic dependent launch (PDL): every kernel is launched with programmaticStreamSerialization so that its CTAs
// are scheduled while the previous kernel of the stream drains; `pdl_prologue()` (griddepcontrol.wait) must run before
// the first access to memory written by earlier kernels, and EVERY kernel must execute it (completion is transitive
// only then).  launch_dependents right away: the next kernel may start its own prologue as soon as SM resources free.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// MUFU ops in their .ftz form: the default (non-ftz) ex2.approx / division intrinsics expand to FSETP + two predicated
// FMUL range fix-ups around every MUFU (ncu source view of the GEGLU GEMM: 36 warp instructions per output element,
// 27 % of them FMUL).  Denormal inputs / results flushed to zero are far below the fp16 store.
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x * sigmoid(x): 2 MUFU + 3 FP (the plain `x / (1 + __expf(-x))` is a full-precision division: ~10 instructions)
// x sigmoid(x) = h + h tanh(h), h = x / 2: ONE MUFU operation (tanh.approx.f32, relative error 2^-11 - the precision of
// the fp16 value it is stored as) instead of the two of x / (1 + 2^(-x log2 e)).  The GroupNorm + SiLU passes are bound
// by the MUFU pipe (4 results per clock and SM sub-partition): level 0, 10.5 M elements = 4.7 us of MUFU time with two.
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// exact-erf GELU (diffusers GEGLU uses F.gelu, approximate="none"): gelu(x) = max(x, 0) - |x (erfc(z) / 2)|, z = |x| / sqrt 2.
// erfc(z) / 2 = 2^q(z) with q a degree-7 polynomial fitted to log2(erfc(z)) - 1 on [0, 4.3] (Chebyshev least squares; beyond
// 4.3 erfc < 1.2e-9 and z is clamped): max |error| of gelu 5.9e-7 over all x with the polynomial evaluated in fp32 - far below
// the fp16 store - and ONE MUFU operation (ex2).  The GEGLU epilogue of the level-0 feed-forward GEMM is bound by the MUFU pipe
// (4 results per clock and SM sub-partition, two epilogue warps on each): the Abramowitz-Stegun 7.1.26 form used before needs
// a reciprocal as well (2 MUFU + 13 FP instructions; erff is ~35).
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752f, 4.3f);
  float q = fmaf(-2.120866156e-05f, z, 4.990906455e-04f);
  q = fmaf(q, z, -5.298239645e-03f);
  q = fmaf(q, z, 3.412367031e-02f);
  q = fmaf(q, z, -1.527471244e-01f);
  q = fmaf(q, z, -9.168493152e-01f);
  q = fmaf(q, z, -1.628133416e+00f);
  q = fmaf(q, z, -9.999946356e-01f);
  const float r = x * mufu_ex2(q);  // x erfc(|x| / sqrt 2) / 2
  return fmaxf(x, 0.0f) - fabsf(r);
}

}  // namespace ls
