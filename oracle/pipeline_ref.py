"""TEST INFRASTRUCTURE - CPU restatement of the pieces of the hot path that live in diffusers==0.32.2 (not vendored
under /root/reference, requirements.txt:4) and of the loop that drives them:

  * DDIMScheduler  (config: /root/reference/configs/scheduler_config.json:1-12; call sites lipsync_pipeline.py:478,562)
  * AutoencoderKL.decode, sd-vae-ft-mse layout (call site lipsync_pipeline.py:145-149; scripts/inference.py:56-58)
  * AutoencoderKL.encode + DiagonalGaussianDistribution.sample (call sites lipsync_pipeline.py:298,315) and the
    prepare_mask_latents / prepare_image_latents helpers around them (:284-320)
  * the segment loop lipsync_pipeline.py:500-575 (the real __call__ cannot run offline: face_alignment on "cuda",
    decord, soundfile, ffmpeg)

PARITY UNPINNED for the two diffusers pieces: the reference ships no tests, golden vectors or fixtures for them
(SURVEY.md §4/§8c).  They restate the published diffusers algorithms; tests pin them to closed forms instead
(abar known answers, F.group_norm / F.conv2d / softmax attention in fp32-fp64).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

SCHEDULER_CONFIG = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                        set_alpha_to_one=False, steps_offset=1, clip_sample=False)


class DDIMRef:
    """diffusers DDIMScheduler, epsilon prediction, 'leading' spacing, eta = 0"""

    def __init__(self, cfg: dict = SCHEDULER_CONFIG):
        self.T = cfg["num_train_timesteps"]
        betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, self.T, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.steps_offset = cfg["steps_offset"]
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n: int):
        self.n = n
        ratio = self.T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = [int(t) for t in ts]
        return self.timesteps

    def step(self, eps: torch.Tensor, t: int, x: torch.Tensor) -> torch.Tensor:
        prev = t - self.T // self.n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


# ----------------------------------------------------------------------------------------------------- VAE decoder
def _gn(sd, p, x, silu):
    y = F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)
    return F.silu(y) if silu else y


def _conv(sd, p, x, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def _resnet2d(sd, p, x):
    """diffusers ResnetBlock2D (temb=None, output_scale_factor=1): GN->SiLU->conv->GN->SiLU->conv, 1x1 shortcut"""
    h = _conv(sd, p + ".conv1", _gn(sd, p + ".norm1", x, True))
    h = _conv(sd, p + ".conv2", _gn(sd, p + ".norm2", h, True))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def _vae_attention(sd, p, x):
    """diffusers Attention inside UNetMidBlock2D (1 head, residual_connection=True, rescale_output_factor=1)"""
    n, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x, False).reshape(n, c, h * w).transpose(1, 2)
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1) @ v
    a = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return a.transpose(1, 2).reshape(n, c, h, w) + x


@torch.no_grad()
def vae_decode(sd: SD, z: torch.Tensor, block_out_channels=(128, 256, 512, 512), layers_per_block=2) -> torch.Tensor:
    """AutoencoderKL.decode(z).sample: post_quant_conv -> Decoder(conv_in, mid[Res, Attn, Res], up blocks, GN-SiLU-conv)"""
    x = _conv(sd, "post_quant_conv", z, padding=0)
    x = _conv(sd, "decoder.conv_in", x)
    x = _resnet2d(sd, "decoder.mid_block.resnets.0", x)
    x = _vae_attention(sd, "decoder.mid_block.attentions.0", x)
    x = _resnet2d(sd, "decoder.mid_block.resnets.1", x)
    n = len(block_out_channels)
    for i in range(n):
        for j in range(layers_per_block + 1):
            x = _resnet2d(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    return _conv(sd, "decoder.conv_out", _gn(sd, "decoder.conv_norm_out", x, True))


# ----------------------------------------------------------------------------------------------------- VAE encoder
@torch.no_grad()
def vae_encode_moments(sd: SD, x: torch.Tensor, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
    """AutoencoderKL.encode(x).latent_dist.parameters: Encoder(conv_in, DownEncoderBlock2D x4, mid[Res, Attn, Res],
    GN-SiLU-conv_out) -> quant_conv.  diffusers Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then 3x3 stride 2.
    Returns (n, 2*latent, h/8, w/8) = [mean | logvar].  PARITY UNPINNED like the decoder (diffusers not vendored)."""
    h = _conv(sd, "encoder.conv_in", x)
    n = len(block_out_channels)
    for i in range(n):
        for j in range(layers_per_block):
            h = _resnet2d(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i != n - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = _resnet2d(sd, "encoder.mid_block.resnets.0", h)
    h = _vae_attention(sd, "encoder.mid_block.attentions.0", h)
    h = _resnet2d(sd, "encoder.mid_block.resnets.1", h)
    h = _conv(sd, "encoder.conv_out", _gn(sd, "encoder.conv_norm_out", h, True))
    return _conv(sd, "quant_conv", h, padding=0)


def gaussian_sample(moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """diffusers DiagonalGaussianDistribution.sample: mean, logvar = chunk(2, dim=1); logvar clamped to [-30, 20];
    mean + exp(0.5 * logvar) * noise (the noise is `randn_tensor(mean.shape, generator=...)` in the library)"""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


@torch.no_grad()
def prepare_image_latents(sd: SD, images: torch.Tensor, noise: torch.Tensor, scaling_factor: float = 0.18215,
                          shift_factor: float = 0.0) -> torch.Tensor:
    """lipsync_pipeline.py:313-320 without the CFG duplicate: (f,3,H,W) -> (1,4,f,h,w)"""
    z = (gaussian_sample(vae_encode_moments(sd, images), noise) - shift_factor) * scaling_factor
    return z.permute(1, 0, 2, 3).unsqueeze(0)


@torch.no_grad()
def prepare_mask_latents(sd: SD, mask: torch.Tensor, masked_image: torch.Tensor, noise: torch.Tensor, height: int,
                         width: int, scaling_factor: float = 0.18215, shift_factor: float = 0.0):
    """lipsync_pipeline.py:284-311 without the CFG duplicate: nearest-resized mask (1,1,f,h,w), masked latents"""
    m = F.interpolate(mask, size=(height // 8, width // 8))
    return m.permute(1, 0, 2, 3).unsqueeze(0), prepare_image_latents(sd, masked_image, noise, scaling_factor,
                                                                     shift_factor)


# ----------------------------------------------------------------------------------------------------- the loop
@torch.no_grad()
def denoise_segment(unet_fn: Callable, seg: Dict[str, torch.Tensor], steps: int = 20, guidance: float = 1.5,
                    trace: Optional[Dict[str, List[torch.Tensor]]] = None, max_steps: Optional[int] = None):
    """lipsync_pipeline.py:537-568 with `unet_fn(sample (B,13,f,h,w), t, audio (B,f,S,D)) -> (B,4,f,h,w)`"""
    sch = DDIMRef()
    ts = sch.set_timesteps(steps)
    lat = seg["latents"].clone()
    do_cfg = guidance > 1.0
    audio = seg["audio_embeds"][None]
    mask, masked, ref = seg["mask_latents"], seg["masked_image_latents"], seg["ref_latents"]
    if do_cfg:
        audio = torch.cat([torch.zeros_like(audio), audio])
        mask, masked, ref = torch.cat([mask] * 2), torch.cat([masked] * 2), torch.cat([ref] * 2)
    for j, t in enumerate(ts):
        if max_steps is not None and j >= max_steps:
            break
        if trace is not None:
            trace.setdefault("latents_in", []).append(lat.clone())
        x = torch.cat([lat] * 2) if do_cfg else lat
        x = torch.cat([x, mask, masked, ref], dim=1)
        eps = unet_fn(x, t, audio)
        if do_cfg:
            eu, ec = eps.chunk(2)
            eps = eu + guidance * (ec - eu)
        lat = sch.step(eps, t, lat)
        if trace is not None:
            trace.setdefault("noise_pred", []).append(eps.clone())
            trace.setdefault("latents", []).append(lat.clone())
    return lat


@torch.no_grad()
def decode_and_paste(vae_sd: SD, lat: torch.Tensor, seg: Dict[str, torch.Tensor], scaling_factor: float = 0.18215):
    """decode_latents (:145-149) + paste_surrounding_pixels_back(decoded, ref, 1 - masks) (:328-333,:572-574)"""
    z = (lat / scaling_factor)[0].permute(1, 0, 2, 3)
    dec = vae_decode(vae_sd, z)
    m = 1 - seg["masks"]
    return dec * m + seg["ref_pixel_values"] * (1 - m)


# ------------------------------------------------------------------------------------- pixel-space pre / post
def preprocess_fixed_mask(images_u8: torch.Tensor, mask_image: torch.Tensor):
    """ImageProcessor.prepare_masks_and_masked_images, fix_mask, affine_transform=False, faces already at the working
    resolution so that transforms.Resize is the identity (image_processor.py:145-165): per frame
    pixel = Normalize([0.5], [0.5])(image / 255.0), masked = pixel * mask_image, mask = mask_image[0:1].
    images_u8: (f,3,H,W) or (f,H,W,3) uint8; mask_image (3,H,W) (float64 in the reference: cv2 resize / 255.0)."""
    if images_u8.shape[3] == 3 and images_u8.shape[1] != 3:
        images_u8 = images_u8.permute(0, 3, 1, 2)
    px, masked, masks = [], [], []
    for img in images_u8:
        p = (img.to(torch.float32) / 255.0 - 0.5) / 0.5
        px.append(p)
        masked.append(p * mask_image)
        masks.append(mask_image[0:1])
    return torch.stack(px), torch.stack(masked), torch.stack(masks)


def restore_faces_u8(faces: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """front half of LipsyncPipeline.restore_video per face (lipsync_pipeline.py:350-355):
    torchvision.transforms.functional.resize(face, (h, w), antialias=True) - which is aten's bilinear anti-aliased
    interpolate, align_corners=False - then "c h w -> h w c", (x / 2 + 0.5).clamp(0, 1) * 255 -> uint8."""
    out = []
    for face in faces:
        r = F.interpolate(face[None].float(), size=(height, width), mode="bilinear", align_corners=False,
                          antialias=True)[0]
        r = (r.permute(1, 2, 0) / 2 + 0.5).clamp(0, 1)
        out.append((r * 255).to(torch.uint8))
    return torch.stack(out)


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    """frames live in [-1, 1] => peak-to-peak 2"""
    mse = (a.double() - b.double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
