/*
 * latentsync_b200 C-ABI — hand-written sm_100a kernels for the LatentSync inference hot path
 * (audio-conditioned UNet3DConditionModel denoising loop + AutoencoderKL decode).
 *
 * The reference (Saltfish-AB/LatentSync) has NO native ABI on this path: everything is nn.Module calls into
 * ATen/cuDNN.  Each entry point below therefore replaces one *library call class* the reference
 * reaches from the cited Python line; the Python host (latentsync_b200/unet.py, vae.py, pipeline.py) keeps the
 * reference's module/pipeline signatures and calls these through ctypes.
 *
 * Conventions
 *  - All pointers passed in are DEVICE pointers owned by the caller (PyTorch); the library never frees or keeps them.
 *    The library itself owns three small per-device scratch areas, allocated with cudaMalloc on FIRST use and kept for
 *    the life of the process: split-K fp32 partials + per-tile tickets (ls_gemm, 64 MB + 32 KB) and GroupNorm partial
 *    sums + tickets (ls_groupnorm*, 4 MB + 128 KB).  They can only grow OUTSIDE stream capture (a call that would
 *    need to grow during capture fails with an error: run the launch sequence once eagerly first, as
 *    engine.Plan.capture does), that growth synchronises the device once, and a superseded block is never freed
 *    because CUDA graphs captured earlier still point at it.  Kernels whose CTAs wait for one another (the
 *    single-launch GroupNorm, split-K GEMMs) are launched cooperatively, so co-residency is enforced by the driver.
 *  - Activations are channels-last fp16: a (b f) x H x W x C video tensor is a row-major [rows, C] matrix with
 *    rows = (b f) * H * W ("tokens").  Weights are packed fp16 [N][K] (K contiguous).
 *  - Every function takes the CUDA stream to launch on (as void*) and returns 0 on success; on failure a
 *    message is available from ls_last_error().  No function synchronises the device (except the one-time scratch
 *    growth above).
 *  - Not re-entrant per stream; one host thread per GPU (the reference's server is single-consumer,
 *    scripts/api.py:24-27,95).
 */
#ifndef LATENTSYNC_B200_H
#define LATENTSYNC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: LsGemmArgs grew (gn_partials_out / gn_unit / gn_partials_ld, up2, stride2 / stride2_pad - all 0 = the version-1
 * behaviour); ls_groupnorm_parts, ls_im2col1d, ls_whisper_chunks added */
#define LS_ABI_VERSION 2

/* returns the message of the last failing call on this thread ("" if none) */
const char* ls_last_error(void);
int ls_abi_version(void);
/* number of kernels launched by this library since process start / last reset (bench.py "gpu_launches") */
int64_t ls_launch_count(void);
void ls_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05.mma + TMEM accumulators + TMA operand staging).
 *
 *   out[m, n] = epilogue( sum_seg sum_tap sum_c  A_seg[pixel(m) + tap, c] * W[n, k(seg, tap, c)] )
 *
 * Replaces: nn.Conv2d 3x3 / 1x1 under InflatedConv3d.forward (latentsync/models/resnet.py:10-18), nn.Linear
 * in Attention / FeedForward / TemporalTransformer3DModel (attention.py:230-235, motion_module.py:102,124),
 * Conv2d proj_in/proj_out (attention.py:55,80), conv_shortcut (resnet.py:180) — the shortcut and the
 * skip-concat (unet_blocks.py:624,745) are expressed as extra K segments of the same GEMM.
 * ------------------------------------------------------------------------------------------------------- */
#define LS_GEMM_MAX_SEG 3
#define LS_EPI_GEGLU 1   /* W rows are packed [value | gate] per N tile; out[m, j] = v * gelu_erf(g) (diffusers GEGLU) */
#define LS_EPI_OUT_F32 2 /* store fp32 instead of fp16 */
#define LS_EPI_SILU 4    /* out = silu(out) */
/* timing probes, honoured only by the -DLS_GEMM_ABLATE build (make gemm_ablate; tools/gemm_ablate.py): the main loop
 * skips the A loads / B loads / MMA issue and the OUTPUT IS GARBAGE; the product build ignores them */
#define LS_DBG_NO_A 256
#define LS_DBG_NO_B 512
#define LS_DBG_NO_MMA 1024

typedef struct LsGemmArgs {
  /* A operand: up to 3 K-segments, each an fp16 channels-last tensor [nimg, H, W, ld] using the first `ch`
   * channels (ch % 64 == 0), with taps = 1 (pointwise) or 9 (3x3, stride 1, zero pad 1). */
  int32_t nseg;
  const void* a_ptr[LS_GEMM_MAX_SEG];
  int32_t a_ch[LS_GEMM_MAX_SEG];
  int32_t a_ld[LS_GEMM_MAX_SEG];
  int32_t a_taps[LS_GEMM_MAX_SEG];
  /* geometry shared by all segments.  A plain [M, K] matrix is nimg = 1, H = 1, W = M.  A batched GEMM over
   * `nimg` problems of M_b rows is nimg = batch, H = 1, W = M_b. */
  int32_t nimg, H, W;
  /* B operand: fp16 [N][Ktot], Ktot = sum_seg taps*ch, K index = ((seg, tap, c)); b_batch_stride != 0 selects a
   * different [N][Ktot] matrix per image (batched GEMM), in elements. */
  const void* b_ptr;
  int32_t N;
  int32_t Ktot;
  int64_t b_batch_stride;
  /* epilogue */
  const float* bias; /* fp32 [bias_rows][N] or NULL */
  int32_t bias_div;  /* rows of `out` per bias row (e.g. F*H*W for a per-batch-element time embedding); 0 = one row */
  int32_t bias_ld;   /* stride between bias rows in floats; 0 = N */
  const void* residual; /* fp16 [M][ldr] or NULL, added after bias */
  int32_t ldr;
  void* out; /* fp16 (or fp32 with LS_EPI_OUT_F32) [M][ldo]; with LS_EPI_GEGLU N_out = N/2 */
  int32_t ldo;
  int32_t flags;
  int32_t tile_n; /* 0 = auto (wave-quantisation cost model); else a multiple of 32 (64 with GEGLU) up to 256 */
  int32_t cta_pair; /* 0 = auto; 1 = single-CTA 128 x tile_n tiles; 2 = CTA pairs (cta_group::2), 256 x tile_n tiles */
  /* nn.LayerNorm folded into the nn.Linear that consumes it (attention.py:176-186,196-197; motion_module.py:208-216):
   * with W' = W diag(gamma) as the B operand and the RAW activations x as A,
   *   LN(x) W^T = rstd[m] (x W'^T)[m, n] - rstd[m] mean[m] col_sum[n] + (beta W^T + bias)[n],
   * col_sum[n] = sum_k W'[n, k] (fp32 [N], of the fp16 operand; GEGLU: packed like the weight rows), the last term goes in
   * `bias` (which may still select a row per `bias_div` output rows: the temporal sinusoid table times W^T).  The row
   * statistics cost no pass of their own: the GEMM that PRODUCES x writes, per row, the (sum, sum of squares) of its
   * output values (fp32, taken just before their rounding to fp16) as a few partials,
   *   row_partials_out: fp32 [n_partials_out][partials_out_stride][2], n_partials_out == 3 * ceil(N / tile_n) (checked:
   *   pass tile_n explicitly; three fixed column ranges per N tile, so that the values do not depend on the launch
   *   geometry), part-major so that rows may be written by several launches (stride >= this launch's M);
   * the consuming GEMM (K = the producer's N, one pointwise segment) sums them in part order (deterministic):
   *   row_partials_in / n_partials_in / partials_in_stride: the producer's array, first row = this launch's row 0;
   *   mean = S / K, rstd = 1 / sqrt(Q / K - mean^2 + ln_eps).
   * Both need the staged fp16 epilogue (ldo % 8 == 0, contiguous rows, no fp32 output) and N % 32 == 0; neither uses
   * split-K.  All NULL / 0 = plain GEMM. */
  const float* col_sum;
  const float* row_partials_in;
  int32_t n_partials_in;
  int64_t partials_in_stride;
  float ln_eps;
  float* row_partials_out;
  int32_t n_partials_out;
  int64_t partials_out_stride;
  /* nn.GroupNorm statistics from the epilogue of the GEMM that PRODUCES the normalised tensor (resnet.py:185,207;
   * attention.py:96; motion_module.py:139 - the reference runs a reduction pass over the tensor for each of them):
   *   gn_partials_out: fp32 [M / 128][gn_partials_ld][2]; entry (t, u) = (sum, sum of squares) over output rows
   *   [128 t, 128 t + 128) and output columns [gn_unit u, gn_unit (u + 1)) of the fp16 values this launch stores.
   * gn_unit = the largest column count that divides the channels-per-group of every GroupNorm reading the tensor (also
   * of a concatenation it is part of).  Needs M % 128 == 0, N % gn_unit == 0, the staged fp16 epilogue without SiLU /
   * GEGLU, and a tile width with N % tile_n == 0 and tile_n % gn_unit == 0 (the automatic choice honours this; an
   * explicit tile_n must).  Summation order is fixed and independent of the launch geometry (deterministic, and
   * identical for any split of the rows over several launches).  No split-K.  ls_groupnorm_parts consumes the array. */
  float* gn_partials_out;
  int32_t gn_unit;
  int32_t gn_partials_ld;
  /* Nearest x2 upsample folded into the 3x3 convolution that follows it (Upsample3D, resnet.py:47-75: F.interpolate(scale 2,
   * "nearest") then conv; diffusers Upsample2D in the VAE decoder): output pixel (2y + py, 2x + px) only ever sees the
   * 2 x 2 low-resolution neighbourhood that starts at (y - 1 + py, x - 1 + px), so the convolution over the upsampled
   * image is four convolutions ("phases") over the LOW-resolution image with 2 x 2 kernels whose weights are sums of the
   * 3x3 taps that land on the same source pixel - 4/9 of the multiply-adds and no upsampled copy.  up2 = 1 + 2 py + px
   * selects the phase: segments carry a_taps = 4 (K index = ((seg, a, b, c)), a / b = row / column of the 2 x 2 window),
   * nimg / H / W are the LOW-resolution geometry, `out` is the high-resolution [nimg, 2H, 2W, ldo] tensor and the launch
   * writes its quarter of the pixels.  Zero padding of the upsampled image = zero padding of the low-resolution one.
   * Needs the staged fp16 epilogue, no residual / GEGLU / partials / split-K.  0 = off. */
  int32_t up2;
  /* 3x3 convolution with stride 2 read in place (Downsample3D, resnet.py:78-101: Conv2d(stride=2, padding=1); diffusers
   * Downsample2D of the VAE encoder: F.pad(x, (0, 1, 0, 1)) + Conv2d(stride=2, padding=0)): nimg / H / W are the OUTPUT geometry,
   * the single 9-tap segment points at the [nimg, 2H, 2W, ld] input, whose pixels are fetched with TMA element strides of two -
   * no im2col copy.  stride2_pad = zero padding before the first row / column (1 or 0; the far side is padded as needed). */
  int32_t stride2;
  int32_t stride2_pad;
} LsGemmArgs;

int ls_gemm(const LsGemmArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Normalisation (fp16 in/out, fp32 statistics).
 * ------------------------------------------------------------------------------------------------------- */
/* GroupNorm statistics over `rows_per_inst` consecutive rows x (C/groups) channels of the virtual concatenation
 * [x1 | x2] (x2 may be NULL).  stats: fp32 [ninst][groups][2] = (sum, sum of squares), fully
 * overwritten.  The reduction is deterministic (fixed-order partials, no floating-point atomics); its scratch is owned
 * by the library and sized by the first (non-captured) call.
 * Replaces the reduction half of nn.GroupNorm: resnet.py:140,164 (5-D input: rows_per_inst = F*H*W),
 * attention.py:51, motion_module.py:101 (per frame: rows_per_inst = H*W), unet.py:236. */
int ls_groupnorm_stats(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows, int32_t rows_per_inst,
                       int32_t groups, float* stats, void* stream);
/* y = (x - mean) * rstd * gamma + beta (+ SiLU), y is [rows][c1+c2] fp16.  resnet.py:185-187,207-215. */
int ls_groupnorm_apply(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows, int32_t rows_per_inst,
                       int32_t groups, const float* stats, const float* gamma, const float* beta, float eps,
                       int32_t silu, void* y, void* stream);
/* Fused single-launch form of the two calls above: y = GroupNorm(x) (+ SiLU).  Statistics, the cross-CTA rendezvous and
 * the normalisation happen in one kernel whenever the launch fits the GPU as one co-resident wave (every UNet / VAE
 * shape); otherwise it falls back to stats + apply.  stats_scratch: fp32 [ninst][groups][2] (used by the fallback). */
int ls_groupnorm(const void* x1, int32_t c1, const void* x2, int32_t c2, int64_t rows, int32_t rows_per_inst,
                 int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu, float* stats_scratch,
                 void* y, void* stream);
/* GroupNorm (+ SiLU) whose statistics come from the GEMM(s) that produced x1 / x2 (LsGemmArgs.gn_partials_out: parts =
 * fp32 [rows / 128][ld][2] per `unit` columns, first tile = the tile of row 0 of x): no reduction pass over the tensor, no
 * cross-CTA rendezvous - each CTA sums the partials of its instance's groups in a fixed order and makes one
 * read-modify-write pass.  rows_per_inst % 128 == 0; `unit` divides (c1 + c2) / groups and c1.  parts2 / ld2 belong to x2.
 * Replaces nn.GroupNorm exactly like ls_groupnorm (resnet.py:185-187,207-215; attention.py:96; motion_module.py:139). */
int ls_groupnorm_parts(const void* x1, int32_t c1, const float* parts1, int32_t ld1, const void* x2, int32_t c2,
                       const float* parts2, int32_t ld2, int64_t rows, int32_t rows_per_inst, int32_t groups, int32_t unit,
                       const float* gamma, const float* beta, float eps, int32_t silu, void* y, void* stream);
/* LayerNorm over C (nn.LayerNorm defaults, attention.py:145,157,172; motion_module.py:195,201), optionally adding
 * the temporal sinusoidal table pe[frame][C] (motion_module.py:232-234) with frame = (row / rows_per_frame) % nframes.
 * pe may be NULL. */
int ls_layernorm(const void* x, int64_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                 const float* pe, int32_t rows_per_frame, int32_t nframes, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Attention (flash-style, fp16 operands, fp32 softmax/accumulate): out = softmax(scale * Q K^T) V.
 * Row addressing:  row(batch b, position i) = (b / inner) * outer_stride + (b % inner) * inner_stride + i * seq_stride
 *   spatial self / audio cross attention (attention.py:250-280): inner = 1, outer_stride = S, seq_stride = 1
 *   temporal attention over frames (motion_module.py:262-313):   inner = H*W, outer_stride = F*H*W, inner_stride = 1,
 *                                                               seq_stride = H*W   (no "(b f) s c -> (b s) f c" copy)
 * ------------------------------------------------------------------------------------------------------- */
typedef struct LsAttnArgs {
  const void* q; /* fp16, row stride ldq, head h at column h*head_dim */
  const void* k;
  const void* v;
  void* out;
  int32_t ldq, ldk, ldv, ldo;
  int32_t batch, heads, head_dim; /* head_dim in {40, 80, 160} */
  int32_t sq, skv;
  int32_t q_inner;
  int64_t q_outer_stride, q_inner_stride, q_seq_stride;
  int32_t kv_inner;
  int64_t kv_outer_stride, kv_inner_stride, kv_seq_stride;
  float scale; /* softmax scale (head_dim^-0.5 in the reference, attention.py:271), applied to the fp32 scores */
} LsAttnArgs;
int ls_attention(const LsAttnArgs* args, void* stream);

/* row softmax for the VAE mid-block attention (diffusers AutoencoderKL: 1 head, d = 512, S = 1024 - too wide for
 * the fused kernel above): p = softmax(scale * s) over `cols`; s fp32 (GEMM with LS_EPI_OUT_F32), p fp16. */
int ls_softmax_rows(const float* s, int64_t rows, int32_t cols, float scale, void* p, void* stream);
/* batched fp16 transpose [batch][R][C] -> [batch][C][R] */
int ls_transpose(const void* x, int32_t batch, int32_t R, int32_t C, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Memory-bound steps of the denoising loop (lipsync_pipeline.py:540-575) and layout helpers.
 * ------------------------------------------------------------------------------------------------------- */
/* cat([latents]*2) + cat([x, mask, masked, ref], dim=1) (lipsync_pipeline.py:542-549) -> channels-last fp16
 * [nb*F*H*W][64] (13 real channels, zero padded to 64 so that conv_in's K is a whole TMA box).
 * latents fp32 [1][4][F][H][W]; mask fp32 [1][1][F][H][W]; masked, ref fp32 [1][4][F][H][W]; nb = 2 with CFG. */
int ls_concat13(const float* latents, const float* mask, const float* masked, const float* ref, int32_t nb, int32_t F,
                int32_t HW, void* out, void* stream);
/* CFG combine + DDIM step (eta = 0), lipsync_pipeline.py:557-562 + diffusers DDIMScheduler.step:
 *   eps = eu + g*(ec - eu);  x0 = (x - sqrt(1-a_t)*eps)/sqrt(a_t);  x_prev = sqrt(a_prev)*x0 + sqrt(1-a_prev)*eps
 * eps_cl: fp32 channels-last [nb*F*HW][ld_eps] (conv_out output; uncond rows first); latents fp32 NCFHW, updated
 * in place; eps_out (optional, fp32 NCFHW [1][4][F][H][W]) receives the guided noise for the parity trace. */
int ls_cfg_ddim_step(const float* eps_cl, int32_t ld_eps, int32_t nb, int32_t F, int32_t HW, float guidance,
                     float alpha_t, float alpha_prev, float* latents, float* eps_out, void* stream);
/* fp32 [B][C][F][HW] -> fp16 channels-last [(b f) HW][cpad] (zero padded, scaled); F = 1 is plain NCHW.
 * Used for UNet3DConditionModel.forward's `sample` (unet.py:312) and decode_latents (lipsync_pipeline.py:145-147). */
int ls_ncfhw_to_cl(const float* x, int32_t B, int32_t C, int32_t F, int32_t HW, int32_t cpad, float scale, void* out,
                   void* stream);
/* fp32 channels-last [(b f) HW][ld] -> fp32 [B][C][F][HW] */
int ls_cl_to_ncfhw(const float* x, int32_t ld, int32_t B, int32_t C, int32_t F, int32_t HW, float* out, void* stream);
/* nearest x2 upsample on (H, W), channels-last fp16 (resnet.py:65; diffusers Upsample2D). */
int ls_upsample2x(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, void* y, void* stream);
/* explicit im2col for the 3x3 stride-2 pad-1 Downsample3D conv (resnet.py:89): out [nimg*(H/2)*(W/2)][9*C]. */
int ls_im2col_s2(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, void* y, void* stream);
/* same with `pad_before` in {0, 1} zero rows/columns before the image (the window always extends to 2*(H/2)+1-pad):
 * pad_before = 0 is diffusers' Downsample2D(padding=0) of the AutoencoderKL encoder - F.pad(x, (0, 1, 0, 1)) then a
 * 3x3 stride-2 convolution (reached from vae.encode, lipsync_pipeline.py:298,315). */
int ls_im2col_s2_pad(const void* x, int32_t nimg, int32_t H, int32_t W, int32_t C, int32_t pad_before, void* y,
                     void* stream);
/* DiagonalGaussianDistribution.sample() followed by (z - shift_factor) * scaling_factor (lipsync_pipeline.py:298-299,
 * 315-316): moments fp32 channels-last [n*HW][ld] = [mean(C) | logvar(C) | ...], logvar clamped to [-30, 20];
 * noise fp32 [n][C][HW] (NULL = the mode); z fp32 [n][C][HW]. */
int ls_gaussian_sample(const float* moments_cl, int32_t ld, const float* noise, int32_t n, int32_t C, int32_t HW,
                       float shift, float scale, float* z, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Pixel-space pre / post processing around the hot path.
 * ------------------------------------------------------------------------------------------------------- */
/* ImageProcessor.preprocess_fixed_mask_image for faces already at the working resolution (the resize is then the
 * identity; image_processor.py:145-151, called per frame from prepare_masks_and_masked_images :153-165):
 * pixel = (u8 / 255 - 0.5) / 0.5, masked = pixel * mask.  img uint8 [n][H][W][3] (hwc != 0) or [n][3][H][W];
 * mask fp32 [mask_c][H][W], mask_c in {1, 3}; pixel, masked fp32 [n][3][H][W]. */
int ls_preprocess_u8(const void* img, int32_t n, int32_t H, int32_t W, int32_t hwc, const float* mask, int32_t mask_c,
                     float* pixel, float* masked, void* stream);
/* restore_video's per-face front half (lipsync_pipeline.py:350-355): torchvision resize(face, (oh, ow),
 * antialias=True) [bilinear, align_corners = false] -> (x / 2 + 0.5).clamp(0, 1) * 255 -> uint8 (truncation),
 * "c h w -> h w c".  x fp32 [n][3][H][W]; out uint8 [n][oh][ow][3]. */
int ls_resize_aa_u8(const float* x, int32_t n, int32_t H, int32_t W, int32_t oh, int32_t ow, void* out, void* stream);
/* paste_surrounding_pixels_back (lipsync_pipeline.py:328-333, called with 1-masks at :572-574):
 *   out = decoded*(1-m) + ref*m.  decoded_cl: fp32 channels-last [n*HW][ld] (VAE conv_out, 3 real channels);
 *   ref fp32 [n][3][HW]; mask fp32 [n][1][HW]; out fp32 [n][3][HW]. */
int ls_paste_back(const float* decoded_cl, int32_t ld, const float* ref, const float* mask, int32_t n, int32_t HW,
                  float* out, void* stream);
/* ---------------------------------------------------------------------------------------------------------
 * Inverse-affine paste-back of the generated faces into the video frames.
 *
 * Replaces AlignRestore.restore_img (latentsync/utils/affine_transform.py:85-115; called per frame from
 * LipsyncPipeline.restore_video, lipsync_pipeline.py:343-358), i.e. OpenCV's warpAffine(INTER_LANCZOS4) of the
 * face, warpAffine of the all-ones mask, erode 2x2, erode (2w x 2w), GaussianBlur(2w+1) and the float blend, for
 * F frames of one size in one call.  Output bytes equal the reference's (OpenCV 4.13 fixed-point arithmetic
 * restated, see csrc/restore.cu).  All pointers are device pointers.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct LsRestoreArgs {
  const void* frames;         /* uint8 [F][H][W][3] video frames */
  void* out;                  /* uint8 [F][H][W][3]; may alias frames (in place) */
  const void* faces;          /* uint8 [F][hf][wf][3] generated faces at the box size (restore_video :350-355) */
  const double* mats;         /* [F][6] dst->src matrices: what cv::warpAffine derives from `inverse_affine` (its
                                 own double-precision inversion), row-major 2x3 */
  const int32_t* rois;        /* [F][4] x0, y0, x1, y1: frame rectangle that contains every pixel the face touches */
  const int16_t* lanczos_tab; /* [32][32][8][8] OpenCV fixed-point Lanczos-4 weights (sum 32768 per entry) */
  const float* gauss_tab;     /* [gmax + 1][2 gmax + 1]: row w = cv2.getGaussianKernel(2 w + 1, 0, CV_32F) */
  float* work;                /* 3 * F * RH * RW floats */
  void* scratch;              /* 12 * F + 4 bytes: uint64 area[F] (sum of the eroded mask * 1024), int32 w_edge[F],
                                 int32 status (set to 1 when some w_edge > gmax: output then invalid) */
  int32_t F, H, W, hf, wf;
  int32_t mh, mw;             /* size of the all-ones mask (AlignRestore.face_size = 280 x 210); 0: same as the face */
  int32_t RW, RH;             /* >= width / height of every ROI */
  int32_t gmax;
} LsRestoreArgs;
int ls_restore_faces(const LsRestoreArgs* args, void* stream);
/* timestep path (unet.py:361-382, resnet.py:190-205): small dense layers on (B, 1280) vectors.
 *   y[b][n] = act_out( sum_k act_in(x[b][k]) * W[n][k] + bias[n] ) (+ add[n]);  x, y, bias, add fp32; W fp16. */
int ls_small_linear(const float* x, int32_t B, int32_t K, const void* W, const float* bias, const float* add,
                    int32_t N, int32_t silu_in, int32_t silu_out, float* y, void* stream);
/* ---------------------------------------------------------------------------------------------------------
 * Whisper-tiny front end: the encoder itself is ls_gemm / ls_layernorm / ls_attention
 * launches (latentsync_b200/whisper.py); these are its two layout helpers.
 * ------------------------------------------------------------------------------------------------------- */
/* nn.Conv1d(kernel 3, padding 1, stride 1 or 2) as a GEMM operand (latentsync/whisper/whisper/model.py:133-134,149-150):
 * x fp16 [n][T][C] -> y fp16 [n][To][3][C], To = (T - 1) / stride + 1 (zeros outside the sequence). */
int ls_im2col1d(const void* x, int32_t n, int32_t T, int32_t C, int32_t stride, void* y, void* stream);
/* Audio2Feature.get_sliced_feature for all video frames in one launch (latentsync/whisper/audio2feature.py:24-48,
 * 85-100): out[i][k * L + l][:] = layers[l][clamp(first[i] + k, 0, T - 1)][:] for k < K, l < L.  layers fp16
 * [L][layer_stride][C] (the embedding and every block output of the encoder), first int32 [n] (device), out fp16 or
 * fp32 (out_f32 != 0) [n][K * L][C]. */
int ls_whisper_chunks(const void* layers, int64_t layer_stride, int32_t L, int32_t T, int32_t C, const int32_t* first,
                      int32_t n, int32_t K, int32_t out_f32, void* out, void* stream);
/* sinusoidal timestep embedding (diffusers get_timestep_embedding, flip_sin_to_cos=True, shift=0): out fp32 [B][dim] */
int ls_timestep_embedding(const float* t, int32_t B, int32_t dim, float* out, void* stream);
int ls_fill_zero(void* p, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LATENTSYNC_B200_H */
