// Memory-bound steps of the denoising loop and layout helpers: 13-channel concat + CFG duplicate, CFG combine +
// DDIM update, NC(F)HW <-> channels-last conversion, nearest x2 upsample, stride-2 im2col, paste-back, the tiny
// timestep-embedding layers.  Single pass, vectorised where the layout allows.
//
// Reference: latentsync/pipelines/lipsync_pipeline.py:542-549 (concat), :557-559 (CFG), :562 + diffusers
// DDIMScheduler.step (update), :328-333,:572-574 (paste-back), latentsync/models/resnet.py:65 (nearest upsample),
// :89 (stride-2 conv), unet.py:361-382 + resnet.py:190-205 (time embedding path).
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

namespace ls {

std::atomic<int64_t> g_launch_count{0};
static thread_local char g_err[512] = {0};

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LS_PDL");  // measured on B200: no gain inside CUDA graphs (15.9 vs 15.7 ms / UNet step)
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void concat13_kernel(const float* __restrict__ lat, const float* __restrict__ mask,
                                const float* __restrict__ masked, const float* __restrict__ ref, int nb, int F, int HW,
                                __half* __restrict__ out) {
  pdl_prologue();
  const int64_t rows = (int64_t)nb * F * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 8) return;
  const int64_t row = idx >> 3;
  const int vec = (int)(idx & 7);
  const int64_t fr = row % ((int64_t)F * HW);  // (f, hw) inside one batch element; both CFG halves share inputs
  const int f = (int)(fr / HW), hw = (int)(fr % HW);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto at = [&](const float* p, int c) { return p[((int64_t)c * F + f) * HW + hw]; };
  if (vec == 0) {
    v[0] = at(lat, 0);
    v[1] = at(lat, 1);
    v[2] = at(lat, 2);
    v[3] = at(lat, 3);
    v[4] = at(mask, 0);
    v[5] = at(masked, 0);
    v[6] = at(masked, 1);
    v[7] = at(masked, 2);
  } else if (vec == 1) {
    v[0] = at(masked, 3);
    v[1] = at(ref, 0);
    v[2] = at(ref, 1);
    v[3] = at(ref, 2);
    v[4] = at(ref, 3);
  }
  uint4 w;
  __half2* o2 = reinterpret_cast<__half2*>(&w);
#pragma unroll
  for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(out + row * 64 + vec * 8) = w;
}

__global__ void cfg_ddim_kernel(const float* __restrict__ eps_cl, int ld, int nb, int F, int HW, float g, float sa_t,
                                float sb_t, float sa_p, float sb_p, float* __restrict__ lat,
                                float* __restrict__ eps_out) {
  pdl_prologue();
  const int64_t n = (int64_t)4 * F * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int c = (int)(idx / ((int64_t)F * HW));
  const int64_t fr = idx % ((int64_t)F * HW);
  const float eu = eps_cl[fr * ld + c];
  float eps = eu;
  if (nb == 2) {
    const float ec = eps_cl[((int64_t)F * HW + fr) * ld + c];
    eps = eu + g * (ec - eu);
  }
  const float x = lat[idx];
  const float x0 = (x - sb_t * eps) / sa_t;
  lat[idx] = sa_p * x0 + sb_p * eps;
  if (eps_out != nullptr) eps_out[idx] = eps;
}

__global__ void ncfhw_to_cl_kernel(const float* __restrict__ x, int B, int C, int F, int HW, int cpad, float scale,
                                   __half* __restrict__ out) {
  pdl_prologue();
  const int64_t rows = (int64_t)B * F * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cpad) return;
  const int64_t row = idx / cpad;
  const int c = (int)(idx % cpad);
  float v = 0.f;
  if (c < C) {
    const int hw = (int)(row % HW);
    const int64_t bf = row / HW;
    const int f = (int)(bf % F), b = (int)(bf / F);
    v = x[(((int64_t)b * C + c) * F + f) * HW + hw] * scale;
  }
  out[idx] = __float2half_rn(v);
}

__global__ void cl_to_ncfhw_kernel(const float* __restrict__ x, int ld, int B, int C, int F, int HW,
                                   float* __restrict__ out) {
  pdl_prologue();
  const int64_t n = (int64_t)B * C * F * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int hw = (int)(idx % HW);
  int64_t r = idx / HW;
  const int f = (int)(r % F);
  r /= F;
  const int c = (int)(r % C);
  const int b = (int)(r / C);
  out[idx] = x[(((int64_t)b * F + f) * HW + hw) * ld + c];
}

__global__ void upsample2x_kernel(const __half* __restrict__ x, int nimg, int H, int W, int C, __half* __restrict__ y) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int64_t total = (int64_t)nimg * (2 * H) * (2 * W) * nvec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    int64_t r = idx / nvec;
    const int wo = (int)(r % (2 * W));
    r /= (2 * W);
    const int ho = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const uint4 u = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + (ho >> 1)) * W + (wo >> 1)) * C + v * 8);
    *reinterpret_cast<uint4*>(y + (((int64_t)n * 2 * H + ho) * 2 * W + wo) * C + v * 8) = u;
  }
}

__global__ void im2col_s2_kernel(const __half* __restrict__ x, int nimg, int H, int W, int C, int pad,
                                 __half* __restrict__ y) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = (int64_t)nimg * Ho * Wo * 9 * nvec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    int64_t r = idx / nvec;
    const int tap = (int)(r % 9);
    r /= 9;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int hi = 2 * ho - pad + tap / 3, wi = 2 * wo - pad + tap % 3;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      u = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + hi) * W + wi) * C + v * 8);
    *reinterpret_cast<uint4*>(y + ((((int64_t)n * Ho + ho) * Wo + wo) * 9 + tap) * C + v * 8) = u;
  }
}

__global__ void paste_back_kernel(const float* __restrict__ dec, int ld, const float* __restrict__ ref,
                                  const float* __restrict__ mask, int n, int HW, float* __restrict__ out) {
  pdl_prologue();
  const int64_t total = (int64_t)n * 3 * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int hw = (int)(idx % HW);
  const int64_t r = idx / HW;
  const int c = (int)(r % 3);
  const int64_t img = r / 3;
  const float m = mask[img * HW + hw];
  const float d = dec[(img * HW + hw) * ld + c];
  out[idx] = d * (1.f - m) + ref[idx] * m;
}

// y[b][n] = act_out(sum_k act_in(x[b][k]) * W[n][k] + bias[n]) + add[n];  one warp per output column
__global__ void small_linear_kernel(const float* __restrict__ x, int B, int K, const __half* __restrict__ W,
                                    const float* __restrict__ bias, const float* __restrict__ add, int N, int silu_in,
                                    int silu_out, float* __restrict__ y) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  for (int b0 = 0; b0 < B; b0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(W + (int64_t)n * K + k);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      float w[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        w[2 * e] = f.x;
        w[2 * e + 1] = f.y;
      }
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        if (b0 + bb < B) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float xv = x[(int64_t)(b0 + bb) * K + k + e];
            if (silu_in) xv = silu_f(xv);
            acc[bb] += xv * w[e];
          }
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[bb] += __shfl_xor_sync(0xffffffffu, acc[bb], o);
      if (lane == 0 && b0 + bb < B) {
        float r = acc[bb] + (bias ? bias[n] : 0.f);
        if (silu_out) r = silu_f(r);
        if (add) r += add[n];
        y[(int64_t)(b0 + bb) * N + n] = r;
      }
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim, float* __restrict__ out) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, i = idx % dim;
  const int half = dim / 2;
  const int k = (i < half) ? i : i - half;
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
  const float a = t[b] * freq;
  out[idx] = (i < half) ? cosf(a) : sinf(a);  // flip_sin_to_cos=True: [cos | sin]
}

}  // namespace ls

using namespace ls;

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
#define LS_LAUNCHED()                                       \
  do {                                                      \
    LS_CUDA(cudaGetLastError());                            \
    g_launch_count.fetch_add(1, std::memory_order_relaxed); \
  } while (0)

extern "C" const char* ls_last_error(void) { return ls::g_err; }
extern "C" int ls_abi_version(void) { return LS_ABI_VERSION; }
extern "C" int64_t ls_launch_count(void) { return g_launch_count.load(); }
extern "C" void ls_reset_launch_count(void) { g_launch_count.store(0); }

extern "C" int ls_concat13(const float* latents, const float* mask, const float* masked, const float* ref, int32_t nb,
                           int32_t F, int32_t HW, void* out, void* stream) {
  LS_CHECK(latents && mask && masked && ref && out && nb >= 1 && nb <= 2 && F > 0 && HW > 0, "ls_concat13: bad args");
  const int64_t n = (int64_t)nb * F * HW * 8;
  LS_CUDA(launch_k(concat13_kernel, dim3(blocks_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), latents, mask, masked, ref, nb, F, HW,
                                                                         (__half*)out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_cfg_ddim_step(const float* eps_cl, int32_t ld_eps, int32_t nb, int32_t F, int32_t HW, float guidance,
                                float alpha_t, float alpha_prev, float* latents, float* eps_out, void* stream) {
  LS_CHECK(eps_cl && latents && nb >= 1 && nb <= 2 && ld_eps >= 4, "ls_cfg_ddim_step: bad args");
  LS_CHECK(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f, "ls_cfg_ddim_step: bad alphas");
  const int64_t n = (int64_t)4 * F * HW;
  LS_CUDA(launch_k(cfg_ddim_kernel, dim3(blocks_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), 
      eps_cl, ld_eps, nb, F, HW, guidance, sqrtf(alpha_t), sqrtf(1.f - alpha_t), sqrtf(alpha_prev),
      sqrtf(1.f - alpha_prev), latents, eps_out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_ncfhw_to_cl(const float* x, int32_t B, int32_t C, int32_t F, int32_t HW, int32_t cpad, float scale,
                              void* out, void* stream) {
  LS_CHECK(x && out && B > 0 && C > 0 && F > 0 && HW > 0 && cpad >= C, "ls_ncfhw_to_cl: bad args");
  const int64_t n = (int64_t)B * F * HW * cpad;
  LS_CUDA(launch_k(ncfhw_to_cl_kernel, dim3(blocks_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), x, B, C, F, HW, cpad, scale, (__half*)out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_cl_to_ncfhw(const float* x, int32_t ld, int32_t B, int32_t C, int32_t F, int32_t HW, float* out,
                              void* stream) {
  LS_CHECK(x && out && B > 0 && C > 0 && F > 0 && HW > 0 && ld >= C, "ls_cl_to_ncfhw: bad args");
  const int64_t n = (int64_t)B * C * F * HW;
  LS_CUDA(launch_k(cl_to_ncfhw_kernel, dim3(blocks_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), x, ld, B, C, F, HW, out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_upsample2x(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, void* y, void* stream) {
  LS_CHECK(x && y && C % 8 == 0, "ls_upsample2x: bad args");
  const int64_t n = (int64_t)nimg * 4 * H * W * (C / 8);
  unsigned blocks = blocks_for(n, 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  LS_CUDA(launch_k(upsample2x_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), (const __half*)x, nimg, H, W, C, (__half*)y));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_im2col_s2(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, void* y, void* stream) {
  LS_CHECK(x && y && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "ls_im2col_s2: bad args");
  const int64_t n = (int64_t)nimg * (H / 2) * (W / 2) * 9 * (C / 8);
  unsigned blocks = blocks_for(n, 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  LS_CUDA(launch_k(im2col_s2_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), (const __half*)x, nimg, H, W, C, 1, (__half*)y));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_im2col_s2_pad(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, int32_t pad_before,
                                void* y, void* stream) {
  LS_CHECK(x && y && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && (pad_before == 0 || pad_before == 1),
           "ls_im2col_s2_pad: bad args");
  const int64_t n = (int64_t)nimg * (H / 2) * (W / 2) * 9 * (C / 8);
  unsigned blocks = blocks_for(n, 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  LS_CUDA(launch_k(im2col_s2_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), (const __half*)x, nimg, H, W, C, pad_before, (__half*)y));
  LS_LAUNCHED();
  return 0;
}

// DiagonalGaussianDistribution.sample + (z - shift) * scale (lipsync_pipeline.py:298-299,315-316):
// moments fp32 channels-last [n*HW][ld] = [mean(C) | logvar(C)], noise fp32 [n][C][HW] -> z fp32 [n][C][HW]
__global__ void gaussian_sample_kernel(const float* __restrict__ mom, int ld, const float* __restrict__ noise, int n,
                                       int C, int HW, float shift, float scale, float* __restrict__ z) {
  pdl_prologue();
  const int64_t total = (int64_t)n * C * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int p = (int)(idx % HW);
  const int c = (int)((idx / HW) % C);
  const int i = (int)(idx / ((int64_t)HW * C));
  const float* row = mom + ((int64_t)i * HW + p) * ld;
  const float mean = row[c];
  const float logvar = fminf(fmaxf(row[C + c], -30.f), 20.f);
  const float eps = noise != nullptr ? noise[idx] : 0.f;  // no noise: the distribution's mode
  z[idx] = (mean + expf(0.5f * logvar) * eps - shift) * scale;
}

extern "C" int ls_gaussian_sample(const float* moments_cl, int32_t ld, const float* noise, int32_t n, int32_t C,
                                  int32_t HW, float shift, float scale, float* z, void* stream) {
  LS_CHECK(moments_cl && z && n > 0 && C > 0 && HW > 0 && ld >= 2 * C, "ls_gaussian_sample: bad args");
  const int64_t total = (int64_t)n * C * HW;
  LS_CUDA(launch_k(gaussian_sample_kernel, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream,
                   moments_cl, ld, noise, n, C, HW, shift, scale, z));
  LS_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Pixel-space pre / post processing around the hot path.
// ---------------------------------------------------------------------------------------------------------------
// ImageProcessor.preprocess_fixed_mask_image without the (identity-size) resize (image_processor.py:145-151):
// pixel = (u8 / 255 - 0.5) / 0.5 ; masked = pixel * mask.  img: uint8 [n][H][W][3] (hwc != 0) or [n][3][H][W];
// mask: fp32 [mask_c][H][W] with mask_c in {1, 3}; pixel, masked: fp32 [n][3][H][W].
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, int n, int HW, int hwc,
                                     const float* __restrict__ mask, int mask_c, float* __restrict__ pixel,
                                     float* __restrict__ masked) {
  pdl_prologue();
  const int64_t total = (int64_t)n * 3 * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int p = (int)(idx % HW);
  const int c = (int)((idx / HW) % 3);
  const int64_t i = idx / ((int64_t)3 * HW);
  const uint8_t u = hwc ? img[(i * HW + p) * 3 + c] : img[idx];
  const float x = (float)u / 255.0f;
  const float v = (x - 0.5f) / 0.5f;
  pixel[idx] = v;
  masked[idx] = v * mask[(mask_c == 3 ? c : 0) * (int64_t)HW + p];
}

extern "C" int ls_preprocess_u8(const void* img, int32_t n, int32_t H, int32_t W, int32_t hwc, const float* mask,
                                int32_t mask_c, float* pixel, float* masked, void* stream) {
  LS_CHECK(img && mask && pixel && masked && n > 0 && H > 0 && W > 0 && (mask_c == 1 || mask_c == 3),
           "ls_preprocess_u8: bad args");
  const int64_t total = (int64_t)n * 3 * H * W;
  LS_CUDA(launch_k(preprocess_u8_kernel, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream,
                   (const uint8_t*)img, n, H * W, hwc, mask, mask_c, pixel, masked));
  LS_LAUNCHED();
  return 0;
}

// torchvision resize(face, (oh, ow), antialias=True) [= aten::_upsample_bilinear2d_aa, align_corners = false] followed
// by (x / 2 + 0.5).clamp(0, 1) * 255 -> uint8 and "c h w -> h w c" (lipsync_pipeline.py:350-355).
// BYTE-exact restatement of that operator's CPU kernel for float tensors (oracle/resize_ref.py states the arithmetic and is
// pinned to torch itself): separable triangle filter, horizontal pass first into a float32 value, a pass whose size does
// not change is skipped; window and weights in the operator's own mixed float / double arithmetic, weights normalised
// by a float DIVISION; and the accumulation order of the x86 builds with FMA: first tap a plain product, the next
// 4 floor((n - 1) / 4) taps product and sum rounded separately (gcc vectorises ATen's loop four taps at a time with an
// in-order reduction), the remaining taps fused.  __fmul_rn / __fadd_rn / __fmaf_rn keep nvcc from re-contracting.
__device__ __forceinline__ void aa_window(int i, float scale, int in_size, int& lo, int& size, float& center,
                                          float& invscale) {
  const float support = scale >= 1.0f ? scale : 1.0f;  // (interp_size * 0.5) * scale, interp_size = 2
  center = (float)((double)scale * ((double)i + 0.5));
  invscale = scale >= 1.0f ? (float)(1.0 / (double)scale) : 1.0f;
  const int max_interp = (int)ceilf(support) * 2 + 1;
  lo = max((int)((double)(center - support) + 0.5), 0);
  size = min((int)((double)(center + support) + 0.5), in_size) - lo;
  size = min(max(size, 0), max_interp);
}
__device__ __forceinline__ float aa_weight(int j, int lo, float center, float invscale) {
  const float t = (float)(j + lo) - center;                       // int64 - float -> float in the operator
  const float x = fabsf((float)(((double)t + 0.5) * (double)invscale));
  return x < 1.0f ? (float)(1.0 - (double)x) : 0.0f;
}
// acc <- next tap, in the operator's order (see above): j = tap index >= 1, nv = 4 floor((n - 1) / 4)
__device__ __forceinline__ float aa_acc(float acc, float s, float w, int j, int nv) {
  return j <= nv ? __fadd_rn(acc, __fmul_rn(s, w)) : __fmaf_rn(s, w, acc);
}

constexpr int AA_MAX_TAPS = 17;  // 2 ceil(scale) + 1: down-scales up to 8x

__global__ void resize_aa_u8_kernel(const float* __restrict__ x, int n, int H, int W, int oh, int ow, float sy,
                                    float sx, uint8_t* __restrict__ out) {
  pdl_prologue();
  const int64_t total = (int64_t)n * oh * ow;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % ow);
  const int oy = (int)((idx / ow) % oh);
  const int64_t i = idx / ((int64_t)ow * oh);
  const bool hpass = (ow != W), vpass = (oh != H);
  int xlo = ox, xn = 1, ylo = oy, yn = 1;
  float xc = 0.f, xi = 1.f, yc = 0.f, yi = 1.f;
  float wx[AA_MAX_TAPS], wy[AA_MAX_TAPS];
  wx[0] = wy[0] = 1.0f;
  if (hpass) {
    aa_window(ox, sx, W, xlo, xn, xc, xi);
    xn = min(xn, AA_MAX_TAPS);
    float tot = 0.f;
    for (int j = 0; j < xn; ++j) {
      wx[j] = aa_weight(j, xlo, xc, xi);
      tot = __fadd_rn(tot, wx[j]);
    }
    if (tot != 0.f)
      for (int j = 0; j < xn; ++j) wx[j] = __fdiv_rn(wx[j], tot);
  }
  if (vpass) {
    aa_window(oy, sy, H, ylo, yn, yc, yi);
    yn = min(yn, AA_MAX_TAPS);
    float tot = 0.f;
    for (int j = 0; j < yn; ++j) {
      wy[j] = aa_weight(j, ylo, yc, yi);
      tot = __fadd_rn(tot, wy[j]);
    }
    if (tot != 0.f)
      for (int j = 0; j < yn; ++j) wy[j] = __fdiv_rn(wy[j], tot);
  }
  const int xnv = ((xn - 1) / 4) * 4, ynv = ((yn - 1) / 4) * 4;
  for (int c = 0; c < 3; ++c) {
    const float* plane = x + (i * 3 + c) * (int64_t)H * W;
    float acc = 0.f;
    for (int r = 0; r < yn; ++r) {
      const float* row = plane + (int64_t)(ylo + r) * W + xlo;
      float h = row[0];
      if (hpass) {  // horizontal pass of this input row (the operator's float32 temporary)
        h = __fmul_rn(row[0], wx[0]);
        for (int j = 1; j < xn; ++j) h = aa_acc(h, row[j], wx[j], j, xnv);
      }
      if (!vpass) acc = h;
      else acc = (r == 0) ? __fmul_rn(h, wy[0]) : aa_acc(acc, h, wy[r], r, ynv);
    }
    const float v = __fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(acc, 0.5f), 0.5f), 0.0f), 1.0f), 255.0f);
    out[idx * 3 + c] = (uint8_t)v;  // truncation, as tensor.to(torch.uint8)
  }
}

extern "C" int ls_resize_aa_u8(const float* x, int32_t n, int32_t H, int32_t W, int32_t oh, int32_t ow, void* out,
                               void* stream) {
  LS_CHECK(x && out && n > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "ls_resize_aa_u8: bad args");
  const float sy = (float)H / (float)oh, sx = (float)W / (float)ow;
  LS_CHECK(2.f * ceilf(sy > 1.f ? sy : 1.f) + 1.f <= (float)AA_MAX_TAPS && 2.f * ceilf(sx > 1.f ? sx : 1.f) + 1.f <= (float)AA_MAX_TAPS,
           "ls_resize_aa_u8: down-scale factor beyond %d taps", AA_MAX_TAPS);
  const int64_t total = (int64_t)n * oh * ow;
  LS_CUDA(launch_k(resize_aa_u8_kernel, dim3(blocks_for(total, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, x, n,
                   H, W, oh, ow, sy, sx, (uint8_t*)out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_paste_back(const float* decoded_cl, int32_t ld, const float* ref, const float* mask, int32_t n,
                             int32_t HW, float* out, void* stream) {
  LS_CHECK(decoded_cl && ref && mask && out && ld >= 3, "ls_paste_back: bad args");
  const int64_t total = (int64_t)n * 3 * HW;
  LS_CUDA(launch_k(paste_back_kernel, dim3(blocks_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), decoded_cl, ld, ref, mask, n, HW, out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_small_linear(const float* x, int32_t B, int32_t K, const void* W, const float* bias, const float* add,
                               int32_t N, int32_t silu_in, int32_t silu_out, float* y, void* stream) {
  LS_CHECK(x && W && y && B > 0 && K % 8 == 0 && N > 0, "ls_small_linear: bad args");
  const int wpb = 8;
  LS_CUDA(launch_k(small_linear_kernel, dim3((N + wpb - 1) / wpb), dim3(wpb * 32), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), x, B, K, (const __half*)W, bias, add,
                                                                                   N, silu_in, silu_out, y));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_timestep_embedding(const float* t, int32_t B, int32_t dim, float* out, void* stream) {
  LS_CHECK(t && out && B > 0 && dim % 2 == 0, "ls_timestep_embedding: bad args");
  LS_CUDA(launch_k(timestep_embedding_kernel, dim3(blocks_for((int64_t)B * dim, 128)), dim3(128), (size_t)(0), (cudaStream_t)((cudaStream_t)stream), t, B, dim, out));
  LS_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------- Whisper front end helpers
// nn.Conv1d(kernel_size = 3, padding = 1, stride s) as a GEMM (latentsync/whisper/whisper/model.py:133-134,149-150):
// x fp16 [n][T][C] -> y fp16 [n][To][3][C], To = (T - 1) / s + 1, tap k of output t reads input s t - 1 + k (zero outside)
__global__ void im2col1d_kernel(const __half* __restrict__ x, int n, int T, int C, int stride, int To,
                                __half* __restrict__ y) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int64_t total = (int64_t)n * To * 3 * nvec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    int64_t r = idx / nvec;
    const int k = (int)(r % 3);
    r /= 3;
    const int to = (int)(r % To);
    const int i = (int)(r / To);
    const int ti = stride * to - 1 + k;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (ti >= 0 && ti < T) u = *reinterpret_cast<const uint4*>(x + ((int64_t)i * T + ti) * C + v * 8);
    *reinterpret_cast<uint4*>(y + (((int64_t)i * To + to) * 3 + k) * C + v * 8) = u;
  }
}

extern "C" int ls_im2col1d(const void* x, int32_t n, int32_t T, int32_t C, int32_t stride, void* y, void* stream) {
  LS_CHECK(x && y && n > 0 && T > 0 && C % 8 == 0 && (stride == 1 || stride == 2), "ls_im2col1d: bad args");
  const int To = (T - 1) / stride + 1;
  const int64_t total = (int64_t)n * To * 3 * (C / 8);
  unsigned blocks = blocks_for(total, 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  LS_CUDA(launch_k(im2col1d_kernel, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __half*)x, n, T, C,
                   stride, To, (__half*)y));
  LS_LAUNCHED();
  return 0;
}

// Audio2Feature.get_sliced_feature for every video frame at once (latentsync/whisper/audio2feature.py:24-48,85-100):
// out[i][k * L + l][:] = layers[l][clamp(first[i] + k, 0, T - 1)][:], k < K (the reference's left_idx .. right_idx - 1 with
// both ends clamped), l < L (the encoder's embedding and block outputs).  layers: fp16 [L][layer_stride rows][C];
// out: fp16 or fp32 [n][K * L][C].  first[i] = int(i * 50 / fps) - 2 * audio_feat_length[0] is computed by the caller with
// the reference's own float arithmetic.
template <typename OutT>
__global__ void whisper_chunks_kernel(const __half* __restrict__ layers, int64_t layer_stride, int L, int T, int C,
                                      const int* __restrict__ first, int n, int K, OutT* __restrict__ out) {
  pdl_prologue();
  const int nvec = C >> 3;
  const int64_t total = (int64_t)n * K * L * nvec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    int64_t r = idx / nvec;
    const int l = (int)(r % L);
    r /= L;
    const int k = (int)(r % K);
    const int i = (int)(r / K);
    int t = first[i] + k;
    t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
    const uint4 u = *reinterpret_cast<const uint4*>(layers + ((int64_t)l * layer_stride + t) * C + v * 8);
    OutT* dst = out + (((int64_t)i * K + k) * L + l) * C + v * 8;
    if constexpr (sizeof(OutT) == 2) {
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]), c = __half22float2(h2[2]),
                   d = __half22float2(h2[3]);
      *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(c.x, c.y, d.x, d.y);
    }
  }
}

extern "C" int ls_whisper_chunks(const void* layers, int64_t layer_stride, int32_t L, int32_t T, int32_t C,
                                 const int32_t* first, int32_t n, int32_t K, int32_t out_f32, void* out, void* stream) {
  LS_CHECK(layers && first && out && L > 0 && T > 0 && C % 8 == 0 && n > 0 && K > 0 && layer_stride >= T,
           "ls_whisper_chunks: bad args");
  const int64_t total = (int64_t)n * K * L * (C / 8);
  unsigned blocks = blocks_for(total, 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (out_f32)
    LS_CUDA(launch_k(whisper_chunks_kernel<float>, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream,
                     (const __half*)layers, layer_stride, L, T, C, (const int*)first, n, K, (float*)out));
  else
    LS_CUDA(launch_k(whisper_chunks_kernel<__half>, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream,
                     (const __half*)layers, layer_stride, L, T, C, (const int*)first, n, K, (__half*)out));
  LS_LAUNCHED();
  return 0;
}

extern "C" int ls_fill_zero(void* p, int64_t bytes, void* stream) {
  LS_CHECK(p && bytes >= 0, "ls_fill_zero: bad args");
  LS_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}
