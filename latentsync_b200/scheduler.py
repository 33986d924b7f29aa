"""DDIMScheduler stand-in with the interface LipsyncPipeline touches: `set_timesteps`, `timesteps`,
`scale_model_input`, `step(...).prev_sample`, `init_noise_sigma`, `order`, `config`.

The reference uses diffusers==0.32.2's DDIMScheduler built from configs/scheduler_config.json
(scripts/inference.py:40); diffusers is not vendored, so this restates its published algorithm for that config:
  betas   = linspace(sqrt(beta_start), sqrt(beta_end), T, fp32)^2           ("scaled_linear")
  abar    = cumprod(1 - betas)
  t_i     = (arange(N) * (T // N)).round()[::-1] + steps_offset                 ("leading" spacing)
  prev_t  = t - T // N ;  abar_prev = abar[prev_t] if prev_t >= 0 else final_alpha_cumprod (= abar[0]: set_alpha_to_one=False)
  x0      = (x - sqrt(1-abar_t) eps) / sqrt(abar_t) ;  x_prev = sqrt(abar_prev) x0 + sqrt(1-abar_prev) eps   (eta = 0)
All per-step coefficients are computed on the host once (no device->host sync inside the loop, unlike indexing
`alphas_cumprod` with a device scalar).  The arithmetic of `step` runs in ls_cfg_ddim_step (pointwise.cu).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import _lib as L

# configs/scheduler_config.json:1-12 + diffusers DDIMScheduler defaults for the keys it omits
DEFAULT_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    trained_betas=None,
    clip_sample=False,
    set_alpha_to_one=False,
    steps_offset=1,
    prediction_type="epsilon",
    thresholding=False,
    timestep_spacing="leading",
    rescale_betas_zero_snr=False,
)


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def alphas_cumprod(cfg: dict) -> np.ndarray:
    T = cfg["num_train_timesteps"]
    if cfg.get("trained_betas") is not None:
        betas = torch.tensor(cfg["trained_betas"], dtype=torch.float32)
    elif cfg["beta_schedule"] == "scaled_linear":
        betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, T, dtype=torch.float32) ** 2
    elif cfg["beta_schedule"] == "linear":
        betas = torch.linspace(cfg["beta_start"], cfg["beta_end"], T, dtype=torch.float32)
    else:
        raise NotImplementedError(cfg["beta_schedule"])
    return torch.cumprod(1.0 - betas, dim=0).numpy()


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, **kwargs):
        cfg = dict(DEFAULT_SCHEDULER_CONFIG)
        for k, v in kwargs.items():
            if k in cfg:
                cfg[k] = v
        if cfg["prediction_type"] != "epsilon" or cfg["thresholding"] or cfg["clip_sample"]:
            raise NotImplementedError("only epsilon prediction without clipping/thresholding (the reference's config)")
        if cfg["timestep_spacing"] != "leading" or cfg["rescale_betas_zero_snr"]:
            raise NotImplementedError("only 'leading' timestep spacing (diffusers default used by the reference)")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self.alphas_cumprod = alphas_cumprod(cfg)  # fp32 numpy, host
        self.final_alpha_cumprod = 1.0 if cfg["set_alpha_to_one"] else float(self.alphas_cumprod[0])
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        """reads <path>/scheduler_config.json like DDIMScheduler.from_pretrained("configs") (scripts/inference.py:40)"""
        import os

        f = path if path.endswith(".json") else os.path.join(path, "scheduler_config.json")
        with open(f) as fh:
            cfg = json.load(fh)
        cfg.update(kw)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        T = self._cfg["num_train_timesteps"]
        if num_inference_steps > T:
            raise ValueError(f"num_inference_steps {num_inference_steps} > num_train_timesteps {T}")
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self._cfg["steps_offset"]
        self._host_timesteps: List[int] = [int(t) for t in ts]
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def step_coefficients(self, timestep: int):
        """(abar_t, abar_prev) for one step, on the host"""
        T = self._cfg["num_train_timesteps"]
        prev = timestep - T // self.num_inference_steps
        a_t = float(self.alphas_cumprod[timestep])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else self.final_alpha_cumprod
        return a_t, a_p

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        """x_t -> x_{t-1}.  model_output / sample: (1 or b, 4, f, h, w) CUDA tensors (any float dtype)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is not used by the reference pipeline (eta=0.0)")
        if not sample.is_cuda:
            raise RuntimeError("latentsync_b200 DDIMScheduler.step needs CUDA tensors (no CPU path)")
        t = int(timestep)  # host value; the pipeline iterates host-side ints, not device scalars
        a_t, a_p = self.step_coefficients(t)
        x = sample.to(torch.float32).contiguous().clone()
        e = model_output.to(torch.float32).contiguous()
        b, c, f, h, w = x.shape
        # reuse the fused kernel with nb = 1 (no CFG combine): it wants eps channels-last [rows, ld]
        e_cl = e.permute(0, 2, 3, 4, 1).reshape(b * f * h * w, c).contiguous()
        for i in range(b):
            xi = x[i]
            L.cfg_ddim_step(e_cl[i * f * h * w:(i + 1) * f * h * w], c, 1, f, h * w, 1.0, a_t, a_p, xi, None)
        out = x.to(sample.dtype)
        if not return_dict:
            return (out,)
        return DDIMSchedulerOutput(prev_sample=out)
