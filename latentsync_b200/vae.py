"""Decoder half of diffusers' `AutoencoderKL` (stabilityai/sd-vae-ft-mse) behind the attributes LipsyncPipeline touches:
`vae.config.{scaling_factor, shift_factor, latent_channels, block_out_channels}` and `vae.decode(z).sample`
(lipsync_pipeline.py:145-149, scripts/inference.py:56-58).  `encode` is upstream of the hot path
and is delegated to a wrapped encoder if one is supplied.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib as L
from .engine import VAEDecoderEngine
from .spec import SD_VAE_FT_MSE_CONFIG, vae_decoder_param_spec


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKLDecoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict = SD_VAE_FT_MSE_CONFIG, device="cuda",
                 encoder=None):
        cfg = dict(SD_VAE_FT_MSE_CONFIG)
        cfg.update(config)
        spec = vae_decoder_param_spec(cfg)
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"VAE decoder state_dict is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, shape in spec.items():
            if tuple(state_dict[k].shape) != tuple(shape):
                # pre-0.20 diffusers checkpoints store the mid attention projections as 1x1 convs / other names;
                # only the Linear layout is accepted here
                raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != expected {shape}")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self.device = torch.device(device)
        self.dtype = torch.float16
        self._encoder = encoder
        L.lib()
        self._engine = VAEDecoderEngine({k: state_dict[k] for k in spec}, cfg, self.device)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def engine(self) -> VAEDecoderEngine:
        return self._engine

    def plan(self, nimg: int, h: int, w: int, capture: bool = True):
        p = self._engine.plan(nimg, h, w)
        if capture and p.graph is None:
            p.capture()
        return p

    def encode(self, x):
        if self._encoder is None:
            raise NotImplementedError("VAE encode is upstream of the accelerated path (SURVEY.md §8f-1); pass "
                                      "`encoder=` (e.g. the diffusers AutoencoderKL) to delegate it")
        return self._encoder.encode(x)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z: (n, 4, h, w) ALREADY divided by scaling_factor (decode_latents does that, lipsync_pipeline.py:146)."""
        if not z.is_cuda:
            raise RuntimeError("latentsync_b200 VAE decode needs CUDA tensors (no CPU path)")
        n, c, h, w = z.shape
        plan = self.plan(n, h, w)
        st = torch.cuda.current_stream().cuda_stream
        lib = L.lib()
        zz = z.to(torch.float32).contiguous()
        L._check(lib.ls_ncfhw_to_cl(zz.data_ptr(), n, c, 1, h * w, plan.z_in.cols, 1.0, plan.z_in.ptr, st),
                 "ls_ncfhw_to_cl")
        plan.replay()
        H, W = plan.out_h, plan.out_w
        out = torch.empty(n, self._cfg["out_channels"], H, W, dtype=torch.float32, device=z.device)
        L._check(lib.ls_cl_to_ncfhw(plan.dec_out.ptr, plan.ld_out, n, self._cfg["out_channels"], 1, H * W,
                                    out.data_ptr(), st), "ls_cl_to_ncfhw")
        out = out.to(z.dtype) if z.dtype in (torch.float16, torch.bfloat16) else out
        if not return_dict:
            return (out,)
        return DecoderOutput(sample=out)
