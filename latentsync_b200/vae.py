"""diffusers' `AutoencoderKL` (stabilityai/sd-vae-ft-mse) behind the attributes LipsyncPipeline touches:
`vae.config.{scaling_factor, shift_factor, latent_channels, block_out_channels}`, `vae.decode(z).sample`
(lipsync_pipeline.py:145-149, scripts/inference.py:56-58) and `vae.encode(x).latent_dist.sample(generator)`
(lipsync_pipeline.py:298,315).  The encoder half is built when the state_dict carries the
`encoder.*` / `quant_conv.*` keys; otherwise `encode` is delegated to a wrapped encoder if one is supplied.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib as L
from .engine import VAEDecoderEngine, VAEEncoderEngine
from .spec import SD_VAE_FT_MSE_CONFIG, vae_decoder_param_spec, vae_encoder_param_spec


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution over moments kept as the encoder plan wrote
    them (fp32 channels-last [(n h w), 2C]); `sample` / `mode` run ls_gaussian_sample and return (n, C, h, w)."""

    def __init__(self, moments_cl: torch.Tensor, n: int, channels: int, h: int, w: int, dtype=torch.float32):
        self._m, self.n, self.c, self.h, self.w, self.dtype = moments_cl, n, channels, h, w, dtype

    def _nchw(self, lo: int) -> torch.Tensor:
        return self._m.view(self.n, self.h, self.w, -1)[..., lo:lo + self.c].permute(0, 3, 1, 2).contiguous()

    @property
    def mean(self) -> torch.Tensor:
        return self._nchw(0).to(self.dtype)

    @property
    def logvar(self) -> torch.Tensor:
        return self._nchw(self.c).clamp(-30.0, 20.0).to(self.dtype)

    @property
    def std(self) -> torch.Tensor:
        return torch.exp(0.5 * self.logvar.float()).to(self.dtype)

    def _draw(self, noise: Optional[torch.Tensor], shift: float = 0.0, scale: float = 1.0) -> torch.Tensor:
        z = torch.empty(self.n, self.c, self.h, self.w, dtype=torch.float32, device=self._m.device)
        nz = None if noise is None else noise.to(self._m.device, torch.float32).contiguous()
        L.gaussian_sample(self._m, self._m.shape[1], nz, self.n, self.c, self.h * self.w, shift, scale, z)
        return z

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """mean + std * randn(mean.shape, generator) - the draw diffusers' randn_tensor makes on the parameters' device"""
        noise = torch.randn((self.n, self.c, self.h, self.w), generator=generator, device=self._m.device,
                            dtype=self.dtype)
        return self._draw(noise).to(self.dtype)

    def sample_scaled(self, noise: Optional[torch.Tensor], shift: float, scale: float) -> torch.Tensor:
        """(sample - shift) * scale in the same pass (lipsync_pipeline.py:299,316); fp32"""
        return self._draw(noise, shift, scale)

    def mode(self) -> torch.Tensor:
        return self._draw(None).to(self.dtype)


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


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
        self._enc_engine = None
        espec = vae_encoder_param_spec(cfg)
        if any(k in state_dict for k in espec):
            miss = [k for k in espec if k not in state_dict]
            if miss:
                raise KeyError(f"VAE encoder state_dict is missing {len(miss)} keys, e.g. {miss[:3]}")
            for k, shape in espec.items():
                if tuple(state_dict[k].shape) != tuple(shape):
                    raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != expected {shape}")
            self._enc_engine = VAEEncoderEngine({k: state_dict[k] for k in espec}, cfg, self.device)

    def to(self, *a, **k):
        """The kernel plans live on the device given at construction; a wrapped `encoder=` module (e.g. a diffusers
        AutoencoderKL that only serves encode) follows the pipeline's `.to(device)` like the reference's VAE does."""
        if self._encoder is not None and hasattr(self._encoder, "to"):
            moved = self._encoder.to(*a, **k)
            if moved is not None:
                self._encoder = moved
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

    def encode_plan(self, nimg: int, H: int, W: int, capture: bool = True):
        if self._enc_engine is None:
            raise NotImplementedError("this AutoencoderKL was built from decoder weights only")
        p = self._enc_engine.plan(nimg, H, W)
        if capture and p.graph is None:
            p.capture()
        return p

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x: (n, 3, H, W) pixels in [-1, 1] -> AutoencoderKLOutput(latent_dist) (lipsync_pipeline.py:298,315)"""
        if self._enc_engine is None:
            if self._encoder is None:
                raise NotImplementedError("this AutoencoderKL was built from decoder weights only; pass the full "
                                          "state_dict (encoder.* / quant_conv.*) or `encoder=` to delegate")
            return self._encoder.encode(x)
        if not x.is_cuda:
            raise RuntimeError("latentsync_b200 VAE encode needs CUDA tensors (no CPU path)")
        n, c, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError(f"height and width must be multiples of 8, got {H}x{W}")
        plan = self.encode_plan(n, H, W)
        st = torch.cuda.current_stream().cuda_stream
        lib = L.lib()
        xx = x.to(torch.float32).contiguous()
        L._check(lib.ls_ncfhw_to_cl(xx.data_ptr(), n, c, 1, H * W, plan.x_in.cols, 1.0, plan.x_in.ptr, st),
                 "ls_ncfhw_to_cl")
        plan.replay()
        lat = self._cfg["latent_channels"]
        dist = DiagonalGaussianDistribution(plan.mom_out.tensor().clone(), n, lat, plan.out_h, plan.out_w,
                                            x.dtype if x.dtype in (torch.float16, torch.bfloat16) else torch.float32)
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(latent_dist=dist)

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


AutoencoderKL = AutoencoderKLDecoder  # full name: the class carries the encoder half when its weights are given

