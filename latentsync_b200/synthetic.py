"""Synthetic weights and segment inputs (there are no checkpoints or videos offline: SURVEY.md §8d).

Everything is a pure function of (key name, shape, seed) through numpy's PCG64 `random()` stream, whose output is
specified to be identical on every platform, so the build container (where the golden vectors are produced with the
reference's own modules) and the GPU box (where they are checked) see bit-identical tensors without shipping 5 GB.

Weight distributions (uniform, like PyTorch's default Conv/Linear init, but variance-preserving so that every branch of
the network contributes at O(1) and a wrong kernel cannot hide behind a residual path):
  * conv / linear weights  U(-a, a), a = sqrt(3 / fan_in)   (var = 1 / fan_in)
  * biases                 U(-0.1, 0.1)
  * norm weight / bias     U(0.8, 1.2) / U(-0.1, 0.1)
  * `zero_module` tensors (conv_in, conv_out, motion proj_out: unet.py:92,241, motion_module.py:65-66) get the same
    treatment as any other layer - with their stock all-zero init the output would be identically 0.
  * pos_encoder.pe buffers are the deterministic sinusoid table of motion_module.py:221-230, not random.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

from .spec import SD_VAE_FT_MSE_CONFIG, unet_param_spec, vae_decoder_param_spec, vae_encoder_param_spec


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed & 0xFFFFFFFF, zlib.crc32(name.encode())]))


def uniform(seed: int, name: str, shape, lo: float, hi: float) -> torch.Tensor:
    n = int(np.prod(shape))
    u = _rng(seed, name).random(n, dtype=np.float32)
    return torch.from_numpy(u).mul_(hi - lo).add_(lo).reshape(tuple(shape))


def approx_normal(seed: int, name: str, shape) -> torch.Tensor:
    """zero-mean unit-variance samples as an Irwin-Hall sum of 12 uniforms (float64 adds: exactly reproducible)"""
    n = int(np.prod(shape))
    u = _rng(seed, name).random((12, n))
    return torch.from_numpy((u.sum(axis=0) - 6.0).astype(np.float32)).reshape(tuple(shape))


def sinusoid_pe(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding buffer (motion_module.py:221-230)"""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def _state_dict(spec, seed: int) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape in spec.items():
        leaf = name.rsplit(".", 1)[-1]
        parent = name.rsplit(".", 1)[0]
        if leaf == "pe":
            sd[name] = sinusoid_pe(shape[1], shape[2])
        elif len(shape) == 1:
            is_norm = any(t in parent.rsplit(".", 1)[-1] for t in ("norm", "norms")) or ".norms." in name
            if is_norm and leaf == "weight":
                sd[name] = uniform(seed, name, shape, 0.8, 1.2)
            else:
                sd[name] = uniform(seed, name, shape, -0.1, 0.1)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = math.sqrt(3.0 / fan_in)
            sd[name] = uniform(seed, name, shape, -a, a)
    return sd


def unet_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 CPU state_dict with the reference's key names (spec.unet_param_spec)"""
    return _state_dict(unet_param_spec(cfg), seed)


def vae_decoder_state_dict(cfg: dict = SD_VAE_FT_MSE_CONFIG, seed: int = 0) -> Dict[str, torch.Tensor]:
    return _state_dict(vae_decoder_param_spec(cfg), seed + 7919)


def vae_encoder_state_dict(cfg: dict = SD_VAE_FT_MSE_CONFIG, seed: int = 0) -> Dict[str, torch.Tensor]:
    return _state_dict(vae_encoder_param_spec(cfg), seed + 104729)


def fixed_mask(height: int = 256, width: int = 256) -> torch.Tensor:
    """Stand-in for latentsync/utils/mask.png (image_processor.py:31-36): 1 = keep the original pixel, 0 = region the
    model repaints.  The real PNG masks rows 95-242, cols 9-247 of a 256x256 face crop; same box,
    scaled."""
    m = torch.ones(1, height, width)
    r0, r1 = round(95 * height / 256), round(243 * height / 256)
    c0, c1 = round(9 * width / 256), round(248 * width / 256)
    m[:, r0:r1, c0:c1] = 0.0
    return m


def segment_inputs(seed: int, segment: int = 0, num_frames: int = 16, height: int = 256, width: int = 256,
                   noise_seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Tensors the denoising loop of one 16-frame segment consumes (lipsync_pipeline.py:503-535), fp32 on CPU.

    latents are ONE (1,4,1,h,w) draw repeated over frames and shared by every segment of a clip
    (prepare_latents, lipsync_pipeline.py:182-196): `noise_seed` does not depend on `segment`."""
    h, w = height // 8, width // 8
    tag = f"seg{segment}"
    lat = approx_normal(noise_seed, "latents", (1, 4, 1, h, w)).repeat(1, 1, num_frames, 1, 1).contiguous()
    audio = approx_normal(seed, tag + ".audio", (num_frames, 50, 384))
    masks = fixed_mask(height, width).unsqueeze(0).repeat(num_frames, 1, 1, 1).contiguous()  # (f,1,H,W)
    mask_lat = torch.nn.functional.interpolate(masks, size=(h, w))  # nearest, as prepare_mask_latents :291-293
    mask_lat = mask_lat.permute(1, 0, 2, 3).unsqueeze(0).contiguous()  # (1,1,f,h,w)
    masked_lat = approx_normal(seed, tag + ".masked", (1, 4, num_frames, h, w)) * 0.9
    ref_lat = approx_normal(seed, tag + ".ref", (1, 4, num_frames, h, w)) * 0.9
    ref_px = uniform(seed, tag + ".refpx", (num_frames, 3, height, width), -1.0, 1.0)
    return dict(latents=lat, audio_embeds=audio, mask_latents=mask_lat, masked_image_latents=masked_lat,
                ref_latents=ref_lat, ref_pixel_values=ref_px, masks=masks)


def restore_case(seed: int, H: int, W: int, scale=(1.2, 1.6), origin=(-20.0, 60.0)):
    """deterministic frame / face / affine matrix (numpy PCG64: identical on every host).  The matrix maps frame ->
    face coordinates like AlignRestore.align_warp_face's: `scale` > 1 makes the face smaller than 210 x 280 in the
    frame, `origin` is the range of the frame position (x and y) of the face's top-left corner."""
    rng = np.random.default_rng(seed)
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    face = rng.integers(0, 256, (280, 210, 3), dtype=np.uint8)
    ang = rng.uniform(-0.3, 0.3)
    sc = rng.uniform(*scale)
    R = np.array([[sc * math.cos(ang), -sc * math.sin(ang)], [sc * math.sin(ang), sc * math.cos(ang)]])
    o = np.array([rng.uniform(*origin), rng.uniform(*origin)])
    A = np.concatenate([R, (-R @ o)[:, None]], axis=1)
    return frame, face, A


def whisper_encoder_state_dict(dims: Dict[str, int], seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 CPU state_dict of `Whisper.encoder` with the checkpoint's key names (whisper.encoder_param_spec): same
    distributions as above (LayerNorm weights U(0.8, 1.2)), `positional_embedding` = the sinusoid buffer of
    latentsync/whisper/whisper/model.py:48-55"""
    from .whisper import encoder_param_spec, sinusoids

    sd = {}
    for name, shape in encoder_param_spec(dims).items():
        if name.endswith("positional_embedding"):
            sd[name] = sinusoids(shape[0], shape[1])
        elif len(shape) == 1:
            is_ln_w = name.endswith(("_ln.weight", "ln_post.weight"))
            sd[name] = uniform(seed + 15485863, name, shape, 0.8, 1.2) if is_ln_w else uniform(seed + 15485863, name, shape, -0.1, 0.1)
        else:
            a = math.sqrt(3.0 / int(np.prod(shape[1:])))
            sd[name] = uniform(seed + 15485863, name, shape, -a, a)
    return sd


def mel_like(seed: int, n_mels: int, n_frames: int) -> torch.Tensor:
    """stand-in for a log-mel spectrogram (whisper/audio.py:120-123 maps it into about [-1, 1.5])"""
    return (approx_normal(seed, "mel", (n_mels, n_frames)) * 0.5 + 0.2).clamp(-1.0, 1.6)
