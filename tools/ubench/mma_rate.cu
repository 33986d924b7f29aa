// Micro-benchmark: cost of a stream of small tcgen05.mma instructions (M = 128, K = 16, fp16 -> fp32) issued by one
// thread, for several N, with 1..3 CTAs resident per SM.  Operand contents do not matter.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mma_rate tools/ubench/mma_rate.cu
#include "../../latentsync_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
namespace ls { void set_error(const char*, ...) {} bool pdl_enabled() { return false; } }
using namespace ls;

template <int N, bool MN_B>
__global__ void mma_stream(int iters, int per_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); tc_fence_before(); }
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 16384;
    const uint32_t idesc = (1u << 4) | (MN_B ? (1u << 16) : 0u) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < per_commit; ++k)
        umma_f16_ss(tm, umma_desc_sw128(a + (k & 3) * 32), umma_desc_sw128(b + (k & 3) * 32), idesc, k > 0);
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1u;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 128); }
}

template <int N, bool MN_B>
void run(const char* name, int ctas_per_sm, int per_commit) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 200;
  const size_t smem = ctas_per_sm == 1 ? 200 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 64 * 1024);
  cudaFuncSetAttribute(mma_stream<N, MN_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_stream<N, MN_B><<<148 * ctas_per_sm, 64, smem>>>(iters, per_commit, d);
  cudaDeviceSynchronize();
  mma_stream<N, MN_B><<<148 * ctas_per_sm, 64, smem>>>(iters, per_commit, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d ctas/SM=%d mma/commit=%2d: %7.1f clk per round trip, %6.1f clk per MMA   (%s)\n", name, N,
         ctas_per_sm, per_commit, (double)h / iters, (double)h / iters / per_commit, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int c = 1; c <= 3; ++c) {
    run<64, false>("S-like (K-major B)", c, 3);
    run<48, true>("PV-like (MN-major B)", c, 4);
    run<64, false>("K-major, long stream", c, 32);
    run<128, false>("K-major, long stream", c, 32);
    run<256, false>("K-major, long stream", c, 32);
    run<48, true>("MN-major, long stream", c, 32);
  }
  return 0;
}
