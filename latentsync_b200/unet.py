"""Drop-in for the reference's `UNet3DConditionModel` (latentsync/models/unet.py:39-512).

Same constructor kwargs, `from_config`, `from_pretrained(model_config, ckpt_path, device) -> (unet, global_step)`,
`load_state_dict` tolerance (unet.py:473-492), state_dict key names (1 246 entries for stage2.yaml) and
`forward(sample, timestep, encoder_hidden_states, ...) -> UNet3DConditionOutput(sample=...)` signature, so that
`LipsyncPipeline`, scripts/inference.py:60-66 and scripts/api.py keep working.  Inference only: the forward runs the
sm_100a kernel plan of engine.UNetEngine (fp16 tensor-core operands, fp32 accumulation); there is no autograd and no
PyTorch fallback - on a machine without the CUDA extension or a GPU, forward raises.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib as L
from .engine import UNetEngine
from .spec import UNET_CTOR_DEFAULTS, unet_config, unet_param_spec, validate_unet_config


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class _Config(dict):
    """attribute + key access, like diffusers' FrozenDict"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _set_nested(root: nn.Module, name: str, value, buffer: bool) -> None:
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    if buffer:
        mod.register_buffer(parts[-1], value)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))


def _sinusoid_table(shape) -> torch.Tensor:
    """PositionalEncoding buffer (motion_module.py:221-230): pe[0, p, 2i] = sin(p w_i), pe[0, p, 2i+1] = cos(p w_i)"""
    _, max_len, d_model = shape
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


class UNet3DConditionModel(nn.Module):
    _supports_gradient_checkpointing = False

    def __init__(self, **kwargs):
        super().__init__()
        unknown = set(kwargs) - set(UNET_CTOR_DEFAULTS)
        if unknown:
            raise TypeError(f"UNet3DConditionModel got unexpected arguments {sorted(unknown)}")
        cfg = unet_config(kwargs)
        validate_unet_config(cfg)
        self._cfg = cfg
        self.config = _Config(cfg)
        self.sample_size = cfg["sample_size"]
        self.use_motion_module = cfg["use_motion_module"]
        self.add_audio_layer = cfg["add_audio_layer"]
        # parameter skeleton with the reference's key names; values stay zero-initialised until load_state_dict
        # (the reference default-initialises; inference always loads a checkpoint, scripts/inference.py:60-64)
        for name, shape in unet_param_spec(cfg).items():
            is_pe = name.endswith(".pe")
            # the sinusoid table is computed, not learned: the reference builds it in PositionalEncoding.__init__
            # (motion_module.py:221-230), so a checkpoint without `pe` entries must still work
            _set_nested(self, name, _sinusoid_table(shape) if is_pe else torch.zeros(shape), buffer=is_pe)
        self._engine: Optional[UNetEngine] = None
        self._engine_key = None
        self._out_dtype = torch.float32

    # ---- construction API of the reference -----------------------------------------------------------------
    @classmethod
    def from_config(cls, config: dict, **kwargs):
        cfg = {k: v for k, v in dict(config).items() if k in UNET_CTOR_DEFAULTS}
        cfg.update(kwargs)
        return cls(**cfg)

    @classmethod
    def from_pretrained(cls, model_config: dict, ckpt_path: str, device="cpu"):
        """unet.py:494-512: (model, resume_global_step); checkpoint = {"state_dict": ..., ["global_step": n]}"""
        unet = cls.from_config(model_config).to(device)
        resume_global_step = 0
        if ckpt_path != "":
            ckpt = torch.load(ckpt_path, map_location=device, weights_only=True)
            resume_global_step = ckpt.get("global_step", 0)
            res = unet.load_state_dict(ckpt["state_dict"], strict=False)
            # every parameter of this skeleton starts at ZERO (the reference default-initialises): a key the checkpoint
            # does not carry would silently run as zeros.  `pe` buffers are computed in the constructor and the
            # shape-mismatched conv_in / conv_out / attn2 entries are dropped on purpose (unet.py:473-492).
            missing = [k for k in res.missing_keys if not k.endswith(".pe")]
            dropped = set(getattr(unet, "_dropped_keys", ()))
            hard = [k for k in missing if k not in dropped]
            if hard:
                raise KeyError(f"checkpoint {ckpt_path!r} lacks {len(hard)} UNet tensors (they would run as zeros), "
                               f"e.g. {hard[:4]}")
            if dropped:
                warnings.warn(f"{len(dropped)} checkpoint tensors were dropped for shape mismatch and stay "
                              f"zero-initialised (unet.py:473-492): {sorted(dropped)[:4]} ...")
            del ckpt
        return unet, resume_global_step

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """drops conv_in / conv_out / attn2.to_k / attn2.to_v entries whose shapes disagree with the config
        (unet.py:473-492) before the normal load"""
        state_dict = dict(state_dict)
        dropped = []
        if "conv_in.weight" in state_dict and state_dict["conv_in.weight"].shape[1] != self.config.in_channels:
            dropped += ["conv_in.weight", "conv_in.bias"]
        if "conv_out.weight" in state_dict and state_dict["conv_out.weight"].shape[0] != self.config.out_channels:
            dropped += ["conv_out.weight", "conv_out.bias"]
        for key in [k for k in state_dict if "attn2.to_k." in k or "attn2.to_v." in k]:
            if state_dict[key].shape[1] != self.config.cross_attention_dim:
                dropped.append(key)
        for key in dropped:
            state_dict.pop(key, None)
        self._dropped_keys = dropped
        res = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._engine = None  # repack on next forward
        return res

    def enable_gradient_checkpointing(self):
        raise RuntimeError("latentsync_b200.UNet3DConditionModel is inference-only (no autograd through CUDA plans)")

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        return out

    # ---- engine plumbing --------------------------------------------------------------------------------------
    def engine(self) -> UNetEngine:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("latentsync_b200.UNet3DConditionModel.forward needs the model on a CUDA device "
                               "(there is no CPU path; call .to('cuda'))")
        if self._engine is None:
            L.lib()  # raises if the CUDA extension is missing
            self._engine = UNetEngine({k: v for k, v in self.state_dict().items()}, self._cfg, dev)
        return self._engine

    def plan(self, B: int, F: int, H: int, W: int, S: int, capture: bool = True, uncond_zero: bool = False,
             same_sample: bool = False):
        p = self.engine().plan(B, F, H, W, S, uncond_zero, same_sample)
        if capture and p.graph is None:
            p.capture()
        return p

    # ---- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(
        self,
        sample: torch.Tensor,
        timestep: Union[torch.Tensor, float, int],
        encoder_hidden_states: torch.Tensor = None,
        class_labels: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
        mid_block_additional_residual: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ):
        if class_labels is not None or attention_mask is not None:
            raise NotImplementedError("class_labels / attention_mask are always None on LatentSync's inference path")
        if down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise NotImplementedError("ControlNet residuals are not part of LatentSync's inference path")
        if self.training:
            raise RuntimeError("inference-only module: call .eval() (lipsync_pipeline.py:388-389 does)")
        if sample.dim() != 5:
            raise ValueError(f"sample must be (batch, channel, frames, height, width), got {tuple(sample.shape)}")
        B, Cin, F, H, W = sample.shape
        if Cin != self.config.in_channels:
            raise ValueError(f"sample has {Cin} channels, config.in_channels = {self.config.in_channels}")
        dev = self.device
        ehs = encoder_hidden_states
        S = 0
        if self.add_audio_layer:
            if ehs is None:
                raise ValueError("encoder_hidden_states is required when add_audio_layer=True")
            if ehs.dim() == 4:  # (b, f, s, d) -> (b f, s, d)  attention.py:184-185
                ehs = ehs.reshape(-1, ehs.shape[-2], ehs.shape[-1])
            if ehs.shape[0] != B * F or ehs.shape[-1] != self.config.cross_attention_dim:
                raise ValueError(f"encoder_hidden_states shape {tuple(encoder_hidden_states.shape)} does not match "
                                 f"batch*frames={B * F}, cross_attention_dim={self.config.cross_attention_dim}")
            S = ehs.shape[1]
        plan = self.plan(B, F, H, W, S)
        # timestep: python number, 0-d tensor or (B,) tensor (unet.py:361-374)
        if not torch.is_tensor(timestep):
            t = torch.full((B,), float(timestep), dtype=torch.float32, device=dev)
        else:
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        plan.t_in.tensor().view(-1)[:B].copy_(t)
        x = sample.to(device=dev, dtype=torch.float32).contiguous()
        if self.config.center_input_sample:
            x = 2 * x - 1.0
        st = torch.cuda.current_stream().cuda_stream
        lib = L.lib()
        L._check(lib.ls_ncfhw_to_cl(x.data_ptr(), B, Cin, F, H * W, plan.x_in.cols, 1.0, plan.x_in.ptr, st),
                 "ls_ncfhw_to_cl")
        if S:
            plan.audio_in.tensor()[:, : ehs.shape[-1]].copy_(ehs.reshape(B * F * S, -1))
        plan.run_hoisted()  # time-embedding path + audio K/V projection: not part of the graph (see engine.Plan.hoisted)
        plan.replay()
        out = torch.empty(B, self.config.out_channels, F, H, W, dtype=torch.float32, device=dev)
        L._check(lib.ls_cl_to_ncfhw(plan.eps_out.ptr, plan.eps_out.cols, B, self.config.out_channels, F, H * W,
                                    out.data_ptr(), st), "ls_cl_to_ncfhw")
        out = out.to(sample.dtype) if sample.dtype in (torch.float16, torch.bfloat16) else out
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)
