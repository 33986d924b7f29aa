// Inverse-affine paste-back of the generated faces into the video frames, batched over frames.
//
// Replaces AlignRestore.restore_img (latentsync/utils/affine_transform.py:85-115), which the reference runs per frame on
// the CPU through OpenCV (opencv-python 4.x; 4.13.0 in this image): cv2.warpAffine(face, INTER_LANCZOS4),
// cv2.warpAffine(ones) (bilinear), two cv2.erode, cv2.GaussianBlur and a float blend truncated to uint8.  The kernels
// restate OpenCV's published algorithms so that the result is the same BYTES (tests/test_restore_gpu.py):
//   * source coordinates in fixed point exactly as cv::warpAffine computes them: dst->src matrix in double,
//     X = (rint((M1 y + M2) 1024) + 16 + rint(M0 x 1024)) >> 5, integer part X >> 5, 5-bit fraction X & 31;
//   * Lanczos-4 (8 x 8 taps) with OpenCV's 32 x 32 table of int16 weights (sum 32768 per entry, built on the host by
//     latentsync_b200/restore.py the way initInterTab2D does), int32 accumulation, (sum + 2^14) >> 15, saturate;
//     taps outside the face contribute 0 (BORDER_CONSTANT, value 0);
//   * the all-ones mask warped bilinearly has the closed form sum of in-bounds (1 - fy | fy)(1 - fx | fx);
//   * erode = min over the structuring rectangle, pixels outside the frame ignored; separable;
//   * GaussianBlur (float, BORDER_REFLECT_101): rows accumulate left to right with FMA, columns accumulate
//     k[0] p0 + sum_j k[j] (p[+j] + p[-j]) with FMA - the order of OpenCV's AVX2 row / symmetric column filters, which is
//     what makes the plateau of the soft mask come out as the same float (1.0 or 1.0000001) as on the CPU;
//   * blend m (e r) + (1 - m) u with separately rounded float32 operations (numpy), truncation to uint8.
// Everything but the final blend only exists inside the face's bounding box in the frame (the ROI, computed on the
// host); outside it the output is the input frame, copied with one cudaMemcpyAsync (or nothing when in place).
//
// Memory-bound byte / float work: no tensor cores.  Per 1080p frame: 6.2 MB frame copy + ~12 B per ROI pixel of mask
// planes + the 176 KB face read through L1/L2.
#include "common.cuh"
#include "../../include/latentsync_b200.h"

#include <atomic>
#include <limits.h>

namespace ls {

extern std::atomic<int64_t> g_launch_count;

struct RestoreP {
  const uint8_t* frames;   // [F][H][W][3]
  uint8_t* out;            // [F][H][W][3]
  const uint8_t* faces;    // [F][hf][wf][3]
  const double* mats;      // [F][6] dst -> src
  const int* rois;         // [F][4] x0, y0, x1, y1 (inside the frame)
  const int16_t* ltab;     // [32][32][8][8]
  const float* gtab;       // [gmax + 1][2 gmax + 1]
  float* e2;               // [F][RH][RW]
  float* t0;               // [F][RH][RW]
  float* t1;               // [F][RH][RW]
  unsigned long long* area;  // [F] sum of e2 * 1024
  int* wedge;              // [F]
  int* status;             // [1] != 0: w_edge beyond the Gaussian table
  int F, H, W, hf, wf, mh, mw, RW, RH, gmax;  // mask: ones(mh, mw) (AlignRestore.face_size), face: hf x wf
};

// cv::saturate_cast<int>(double) == cvRound: round half to even
__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }

// fixed-point source coordinate of dst pixel (x, y): integer parts and 5-bit fractions (imgwarp.cpp, WarpAffineInvoker)
__device__ __forceinline__ void src_coord(const double* M, int x, int y, int& sx, int& sy, int& fx, int& fy) {
  const int adelta = cv_round(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = cv_round(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const int X0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
  const int Y0 = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  sx = max(-32768, min(32767, X >> 5));  // saturate_cast<short>
  sy = max(-32768, min(32767, Y >> 5));
  fx = X & 31;
  fy = Y & 31;
}

// cv2.warpAffine(ones(hf, wf) float32, ...) at dst pixel (x, y): in-bounds bilinear weights
__device__ __forceinline__ float mask_at(const double* M, int x, int y, int wf, int hf) {
  int sx, sy, fx, fy;
  src_coord(M, x, y, sx, sy, fx, fy);
  if (sx < -1 || sy < -1 || sx >= wf || sy >= hf) return 0.f;
  const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
  const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty};
  float s = 0.f;
#pragma unroll
  for (int ky = 0; ky < 2; ++ky)
#pragma unroll
    for (int kx = 0; kx < 2; ++kx) {
      const int xx = sx + kx, yy = sy + ky;
      if (xx >= 0 && xx < wf && yy >= 0 && yy < hf) s = __fadd_rn(s, __fmul_rn(wy[ky], wx[kx]));
    }
  return s;
}

// (1) e2 = erode(mask, 2 x 2) (anchor (1, 1): window {y - 1, y} x {x - 1, x}) and the exact sum of e2 per frame.  The
// warped mask is evaluated once per pixel of the block's 33 x 9 footprint (left / upper halo) into shared memory.
__global__ void __launch_bounds__(256) restore_mask_kernel(const RestoreP p) {
  pdl_prologue();
  __shared__ float mk[9][33];
  const int f = blockIdx.z;
  const int* roi = p.rois + 4 * f;
  const int bx = roi[0] + blockIdx.x * 32, by = roi[1] + blockIdx.y * 8;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  double M[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) M[i] = p.mats[6 * f + i];
  for (int i = tid; i < 9 * 33; i += 256) {
    const int ly = i / 33, lx = i - ly * 33;
    const int xx = bx - 1 + lx, yy = by - 1 + ly;
    // outside the frame: ignored by the erosion (+inf); right / lower overhang of the block is never read
    mk[ly][lx] = (xx < 0 || yy < 0) ? INFINITY : ((xx < p.W && yy < p.H) ? mask_at(M, xx, yy, p.mw, p.mh) : 0.f);
  }
  __syncthreads();
  const int rx = blockIdx.x * 32 + threadIdx.x, ry = blockIdx.y * 8 + threadIdx.y;
  const int x = roi[0] + rx, y = roi[1] + ry;
  unsigned int fix = 0;
  if (x < roi[2] && y < roi[3]) {
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    const float e = fminf(fminf(mk[ly - 1][lx - 1], mk[ly - 1][lx]), fminf(mk[ly][lx - 1], mk[ly][lx]));
    p.e2[((size_t)f * p.RH + ry) * p.RW + rx] = e;
    fix = (unsigned int)(e * 1024.0f);  // multiples of 1 / 1024: exact
  }
  // block sum -> ONE atomic per block (one per warp put 8 k same-address atomics on each frame's counter: 68 of the
  // kernel's 74 us)
  __shared__ unsigned int wsum[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fix += __shfl_xor_sync(0xffffffffu, fix, o);
  if (threadIdx.x == 0) wsum[threadIdx.y] = fix;
  __syncthreads();
  if (tid == 0) {
    unsigned int tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += wsum[i];
    if (tot != 0) atomicAdd(p.area + f, (unsigned long long)tot);
  }
}

// (2) w_edge = int(sqrt(float32 sum)) // 20 (affine_transform.py:101-102)
__global__ void restore_wedge_kernel(const RestoreP p) {
  pdl_prologue();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.F) return;
  const float area = (float)((double)p.area[f] * (1.0 / 1024.0));
  const int w = (int)sqrtf(area) / 20;
  p.wedge[f] = w;
  if (w > p.gmax) atomicExch(p.status, 1);
}

// value of a ROI plane at absolute frame position (x, y), y / x inside the frame: 0 outside the ROI
__device__ __forceinline__ float plane_at(const float* pl, const int* roi, int RW, int x, int y) {
  if (x < roi[0] || x >= roi[2] || y < roi[1] || y >= roi[3]) return 0.f;
  return pl[(size_t)(y - roi[1]) * RW + (x - roi[0])];
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

// (3) + (4) the two erosions and the blur are 1-D window passes over a ROI plane.  A block stages its rows / columns
// plus a halo of w on both sides in shared memory (border rules applied while loading), then every output walks its
// window there: no bounds checks and no global loads inside the window loop.
//   erode, ones(2 w, 2 w), anchor (w, w): min over offsets -w .. w - 1, pixels outside the frame ignored (+inf);
//          w == 0: cv2 substitutes a 3 x 3 rectangle for the empty kernel (offsets -1 .. 1)
//   blur,  (2 w + 1) taps, BORDER_REFLECT_101: rows accumulate left to right with FMA, columns accumulate
//          k[w] p0 + sum_j k[w + j] (p[+j] + p[-j]) with FMA (OpenCV's order, see the header)
// Row passes: 128 x 2 outputs per block; column passes: 32 x 32 outputs (four per thread).
constexpr int WIN_HALO = 128;  // == the largest w with a Gaussian kernel (gmax <= WIN_HALO is checked on the host)
constexpr int ROW_TX = 128, ROW_TY = 2, COL_TX = 32, COL_TY = 32;

template <bool ROWS, bool BLUR>
__global__ void __launch_bounds__(256) restore_window_kernel(const RestoreP p, const float* __restrict__ src,
                                                             float* __restrict__ dst) {
  pdl_prologue();
  constexpr int TX = ROWS ? ROW_TX : COL_TX, TY = ROWS ? ROW_TY : COL_TY;
  constexpr int LEN = (ROWS ? TX : TY) + 2 * WIN_HALO;  // window axis, with halo
  constexpr int LINES = ROWS ? TY : TX;                 // independent lines per block
  __shared__ float tile[LINES * LEN];                   // ROWS: [line = y][t along x]; COLS: [t along y][line = x]
  const int f = blockIdx.z;
  const int* roi = p.rois + 4 * f;
  const int w = p.wedge[f];
  if (w > p.gmax || w > WIN_HALO) return;  // flagged by restore_wedge_kernel: the output is invalid anyway
  const int x0 = roi[0] + blockIdx.x * TX, y0 = roi[1] + blockIdx.y * TY;
  if (x0 >= roi[2] || y0 >= roi[3]) return;
  const int hw = (w == 0) ? 1 : w;
  const int n_axis = (ROWS ? TX : TY) + 2 * hw;
  const float* pl = src + (size_t)f * p.RH * p.RW;
  const int tid = threadIdx.x;
  for (int i = tid; i < n_axis * LINES; i += 256) {
    int t, line, x, y;
    if (ROWS) {
      line = i / n_axis;
      t = i - line * n_axis;
      x = x0 - hw + t;
      y = y0 + line;
    } else {
      t = i / LINES;
      line = i - t * LINES;
      x = x0 + line;
      y = y0 - hw + t;
    }
    float v;
    if (BLUR) {
      const int xr = ROWS ? reflect101(x, p.W) : x, yr = ROWS ? y : reflect101(y, p.H);
      v = (xr < p.W && yr < p.H) ? plane_at(pl, roi, p.RW, xr, yr) : 0.f;
    } else {
      v = (x < 0 || y < 0 || x >= p.W || y >= p.H) ? INFINITY : plane_at(pl, roi, p.RW, x, y);
    }
    tile[ROWS ? line * LEN + t : t * LINES + line] = v;
  }
  __syncthreads();
  const float* k = p.gtab + (size_t)w * (2 * p.gmax + 1);
  constexpr int PER = (TX * TY) / 256;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int idx = tid + q * 256;
    const int lx = idx % TX, ly = idx / TX;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= roi[2] || y >= roi[3]) continue;
    const int c = (ROWS ? lx : ly) + hw;                     // centre position along the window axis
    const float* line = ROWS ? tile + ly * LEN : tile + lx;  // element t of the line: line[t * stride]
    constexpr int stride = ROWS ? 1 : LINES;
    float r;
    if (!BLUR) {
      const int lo = (w == 0) ? -1 : -w, hi = (w == 0) ? 1 : w - 1;
      r = INFINITY;
      for (int d = lo; d <= hi; ++d) r = fminf(r, line[(c + d) * stride]);
    } else if (w == 0) {
      r = line[c * stride];  // 1 x 1 kernel
    } else if (ROWS) {
      r = __fmul_rn(line[(c - w) * stride], k[0]);
      for (int j = 1; j <= 2 * w; ++j) r = __fmaf_rn(line[(c - w + j) * stride], k[j], r);
    } else {
      r = __fmul_rn(line[c * stride], k[w]);
      for (int j = 1; j <= w; ++j) r = __fmaf_rn(__fadd_rn(line[(c + j) * stride], line[(c - j) * stride]), k[w + j], r);
    }
    dst[((size_t)f * p.RH + (y - roi[1])) * p.RW + (x - roi[0])] = r;
  }
}

// 8 x 8 Lanczos taps of one pixel from `base` (shared-memory tile or the face in global memory; inlined per call site
// so that the tile reads compile to LDS): row pointer = base + (yy - oy) * pitch, texel xx - ox.  Taps outside the face
// contribute 0; pixels whose 8 x 8 window lies inside the face skip the per-tap checks.
__device__ __forceinline__ void lanczos_acc(const uint8_t* base, int pitch, int ox, int oy, int sx, int sy, int wf, int hf,
                                            const uint4* wq, int (&acc)[3]) {
  const bool interior = sx >= 3 && sy >= 3 && sx + 4 < wf && sy + 4 < hf;
  if (interior) {
#pragma unroll 2
    for (int ky = 0; ky < 8; ++ky) {
      const uint4 w8 = __ldg(wq + ky);
      const int16_t* wv = reinterpret_cast<const int16_t*>(&w8);
      const uint8_t* px = base + (sy - 3 + ky - oy) * pitch + (sx - 3 - ox) * 3;
#pragma unroll
      for (int kx = 0; kx < 8; ++kx) {
        const int w = wv[kx];
        acc[0] += w * px[kx * 3];
        acc[1] += w * px[kx * 3 + 1];
        acc[2] += w * px[kx * 3 + 2];
      }
    }
    return;
  }
  for (int ky = 0; ky < 8; ++ky) {
    const int yy = sy - 3 + ky;
    if (yy < 0 || yy >= hf) continue;
    const uint4 w8 = __ldg(wq + ky);
    const int16_t* wv = reinterpret_cast<const int16_t*>(&w8);
    const uint8_t* row = base + (yy - oy) * pitch - ox * 3;
#pragma unroll
    for (int kx = 0; kx < 8; ++kx) {
      const int xx = sx - 3 + kx;
      if (xx < 0 || xx >= wf) continue;
      const int w = wv[kx];
      acc[0] += w * row[xx * 3];
      acc[1] += w * row[xx * 3 + 1];
      acc[2] += w * row[xx * 3 + 2];
    }
  }
}

// (5) Lanczos-4 warp of the face + blend, ROI pixels only.  A 32 x 8 block of frame pixels reads a small rectangle of
// the face (the block's source bounding box + the 8-tap reach): it is staged in shared memory once (coalesced byte
// rows) and the 192 byte reads per pixel hit shared memory instead of L1 tags; the 64 int16 weights of a pixel are
// eight 16-byte loads.  (First version: 256 scalar global loads per pixel, 949 us for 16 x 1080p frames.)
constexpr int BLEND_SMEM = 40 * 1024;

__global__ void __launch_bounds__(256) restore_blend_kernel(const RestoreP p, const float* __restrict__ soft) {
  pdl_prologue();
  __shared__ __align__(16) uint8_t tile[BLEND_SMEM];
  __shared__ int box[4];  // min sx, min sy, max sx, max sy over the pixels that need the warp
  const int f = blockIdx.z;
  const int* roi = p.rois + 4 * f;
  const int rx = blockIdx.x * 32 + threadIdx.x, ry = blockIdx.y * 8 + threadIdx.y;
  const int x = roi[0] + rx, y = roi[1] + ry;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const bool inside = (x < roi[2] && y < roi[3]);
  if (tid == 0) {
    box[0] = box[1] = INT_MAX;
    box[2] = box[3] = INT_MIN;
  }
  __syncthreads();
  const size_t o = ((size_t)f * p.RH + ry) * p.RW + rx;
  const float m = inside ? soft[o] : 0.f, e = inside ? p.e2[o] : 0.f;
  const bool warp_px = inside && m != 0.f && e != 0.f;  // m == 0: out = frame; e == 0: the face term is 0
  int sx = 0, sy = 0, fx = 0, fy = 0;
  if (warp_px) {
    double M[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) M[i] = p.mats[6 * f + i];
    src_coord(M, x, y, sx, sy, fx, fy);
  }
  {  // warp-level bounding box, then one shared-memory atomic per warp and bound
    const int lo_x = __reduce_min_sync(0xffffffffu, warp_px ? sx : INT_MAX);
    const int lo_y = __reduce_min_sync(0xffffffffu, warp_px ? sy : INT_MAX);
    const int hi_x = __reduce_max_sync(0xffffffffu, warp_px ? sx : INT_MIN);
    const int hi_y = __reduce_max_sync(0xffffffffu, warp_px ? sy : INT_MIN);
    if (threadIdx.x == 0 && lo_x != INT_MAX) {
      atomicMin(&box[0], lo_x);
      atomicMin(&box[1], lo_y);
      atomicMax(&box[2], hi_x);
      atomicMax(&box[3], hi_y);
    }
  }
  __syncthreads();
  const uint8_t* face = p.faces + (size_t)f * p.hf * p.wf * 3;
  // rectangle of the face the block touches, clipped to the face
  const int bx0 = max(0, box[0] - 3), by0 = max(0, box[1] - 3);
  const int bx1 = min(p.wf, box[2] + 5), by1 = min(p.hf, box[3] + 5);
  const int bw = bx1 - bx0, bh = by1 - by0;
  const bool any = box[0] != INT_MAX && bw > 0 && bh > 0;
  const bool staged = any && (size_t)bw * 3 * bh <= BLEND_SMEM;
  if (staged) {
    const int rowb = bw * 3;
    for (int i = tid; i < rowb * bh; i += 256) {
      const int r = i / rowb, c = i - r * rowb;
      tile[i] = face[((size_t)(by0 + r) * p.wf + bx0) * 3 + c];
    }
  }
  __syncthreads();
  if (!inside) return;
  const size_t pix = (((size_t)f * p.H + y) * p.W + x) * 3;
  if (m == 0.f) {  // 0 * (e r) + 1 * u == u
    if (p.out != p.frames) {
      p.out[pix] = p.frames[pix];
      p.out[pix + 1] = p.frames[pix + 1];
      p.out[pix + 2] = p.frames[pix + 2];
    }
    return;
  }
  int acc[3] = {0, 0, 0};
  if (warp_px) {
    const uint4* wq = reinterpret_cast<const uint4*>(p.ltab + (size_t)(fy * 32 + fx) * 64);
    if (staged)
      lanczos_acc(tile, bw * 3, bx0, by0, sx, sy, p.wf, p.hf, wq, acc);
    else
      lanczos_acc(face, p.wf * 3, 0, 0, sx, sy, p.wf, p.hf, wq, acc);
  }
  const float om = __fsub_rn(1.0f, m);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int r8 = max(0, min(255, (acc[c] + (1 << 14)) >> 15));
    const float v = __fadd_rn(__fmul_rn(m, __fmul_rn(e, (float)r8)), __fmul_rn(om, (float)p.frames[pix + c]));
    p.out[pix + c] = (uint8_t)(int)v;  // numpy astype(uint8): truncation
  }
}

static int restore_impl(const LsRestoreArgs* a, cudaStream_t stream) {
  LS_CHECK(a != nullptr, "ls_restore_faces: null args");
  LS_CHECK(a->F >= 1 && a->H >= 1 && a->W >= 1 && a->hf >= 8 && a->wf >= 8, "ls_restore_faces: bad geometry");
  LS_CHECK(a->H < 32768 && a->W < 32768, "ls_restore_faces: frame larger than OpenCV's short coordinates");
  LS_CHECK(a->RW >= 1 && a->RH >= 1 && a->gmax >= 1 && a->gmax <= WIN_HALO, "ls_restore_faces: bad ROI / table size (gmax <= %d)", WIN_HALO);
  LS_CHECK(a->frames && a->out && a->faces && a->mats && a->rois && a->lanczos_tab && a->gauss_tab && a->work &&
               a->scratch,
           "ls_restore_faces: null pointer");
  RestoreP p;
  p.frames = reinterpret_cast<const uint8_t*>(a->frames);
  p.out = reinterpret_cast<uint8_t*>(a->out);
  p.faces = reinterpret_cast<const uint8_t*>(a->faces);
  p.mats = a->mats;
  p.rois = a->rois;
  p.ltab = a->lanczos_tab;
  p.gtab = a->gauss_tab;
  const size_t plane = (size_t)a->F * a->RH * a->RW;
  p.e2 = a->work;
  p.t0 = a->work + plane;
  p.t1 = a->work + 2 * plane;
  // scratch: F x uint64 area | F x int32 w_edge | int32 status
  p.area = reinterpret_cast<unsigned long long*>(a->scratch);
  p.wedge = reinterpret_cast<int*>(p.area + a->F);
  p.status = p.wedge + a->F;
  p.F = a->F;
  p.H = a->H;
  p.W = a->W;
  p.hf = a->hf;
  p.wf = a->wf;
  p.mh = a->mh > 0 ? a->mh : a->hf;
  p.mw = a->mw > 0 ? a->mw : a->wf;
  p.RW = a->RW;
  p.RH = a->RH;
  p.gmax = a->gmax;
  LS_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)a->F * 12 + 4, stream));
  if (p.out != p.frames)
    LS_CUDA(cudaMemcpyAsync(p.out, p.frames, (size_t)a->F * a->H * a->W * 3, cudaMemcpyDeviceToDevice, stream));
  const dim3 block(32, 8), grid((a->RW + 31) / 32, (a->RH + 7) / 8, a->F);
  LS_CUDA(launch_k(restore_mask_kernel, grid, block, (size_t)0, stream, p));
  LS_CUDA(launch_k(restore_wedge_kernel, dim3((a->F + 127) / 128), dim3(128), (size_t)0, stream, p));
  const dim3 wblock(256);
  const dim3 grid_r((a->RW + ROW_TX - 1) / ROW_TX, (a->RH + ROW_TY - 1) / ROW_TY, a->F);
  const dim3 grid_c((a->RW + COL_TX - 1) / COL_TX, (a->RH + COL_TY - 1) / COL_TY, a->F);
  LS_CUDA(launch_k(restore_window_kernel<true, false>, grid_r, wblock, (size_t)0, stream, p, (const float*)p.e2, p.t0));
  LS_CUDA(launch_k(restore_window_kernel<false, false>, grid_c, wblock, (size_t)0, stream, p, (const float*)p.t0, p.t1));
  LS_CUDA(launch_k(restore_window_kernel<true, true>, grid_r, wblock, (size_t)0, stream, p, (const float*)p.t1, p.t0));
  LS_CUDA(launch_k(restore_window_kernel<false, true>, grid_c, wblock, (size_t)0, stream, p, (const float*)p.t0, p.t1));
  LS_CUDA(launch_k(restore_blend_kernel, grid, block, (size_t)0, stream, p, (const float*)p.t1));
  LS_CUDA(cudaGetLastError());
  g_launch_count.fetch_add(7, std::memory_order_relaxed);
  return 0;
}

}  // namespace ls

extern "C" int ls_restore_faces(const LsRestoreArgs* args, void* stream) {
  return ls::restore_impl(args, reinterpret_cast<cudaStream_t>(stream));
}
