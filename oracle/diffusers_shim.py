"""TEST INFRASTRUCTURE - not part of the product (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import anything under oracle/).

Stand-in for the nine `diffusers==0.32.2` symbols the reference's model files import (requirements.txt:4; the package
is NOT vendored under /root/reference and is not installed here), so that the reference's OWN
latentsync/models/{unet,unet_blocks,resnet,attention,motion_module}.py execute unmodified on CPU:

    diffusers.configuration_utils.{ConfigMixin, register_to_config}    unet.py:11, attention.py:10
    diffusers.models.ModelMixin                                          unet.py:12
    diffusers.utils.{BaseOutput, logging}                                unet.py:14
    diffusers.models.embeddings.{TimestepEmbedding, Timesteps}          unet.py:15,95-98
    diffusers.models.attention.{FeedForward, AdaLayerNorm}              attention.py:13,171 ; motion_module.py:16,200
    latentsync.utils.util.zero_rank_log                                  unet.py:27 (real module needs decord/mediapipe)

PARITY UNPINNED at this boundary: the reference holds no tests or golden vectors for these pieces, and
the diffusers source is absent, so Timesteps / TimestepEmbedding / FeedForward(GEGLU) below restate the published
diffusers 0.32.2 algorithms (see oracle/README.md) and are cross-checked against closed forms in tests/.
"""
from __future__ import annotations

import functools
import inspect
import logging as _pylogging
import math
import os
import sys
import types
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("LATENTSYNC_REFERENCE", "/root/reference")


class FrozenDict(OrderedDict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(name) from e


def register_to_config(init):
    """records the constructor arguments (defaults included) in self.config, like diffusers does"""

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        init(self, *args, **kwargs)
        self._internal_dict = FrozenDict(cfg)

    return wrapper


class ConfigMixin:
    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        params = set(inspect.signature(cls.__init__).parameters) - {"self"}
        return cls(**{k: v for k, v in dict(config).items() if k in params}, **kwargs)


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class BaseOutput(OrderedDict):
    """dataclass base with attribute, key and index access"""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)

    def to_tuple(self):
        return tuple(self.values())


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    """diffusers.models.embeddings.get_timestep_embedding (sinusoidal, fp32)"""
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale=1):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift
        self.scale = scale

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift,
                                      self.scale)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu"):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu"):
        super().__init__()
        assert activation_fn == "geglu"
        inner = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out)])

    def forward(self, hidden_states):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class AdaLayerNorm(nn.Module):  # imported by attention.py:13, never constructed on this path
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaLayerNorm is not used by LatentSync's inference path")


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install() -> None:
    """register the stand-ins and put the reference tree on sys.path (idempotent)"""
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_latentsync_b200_shim", False):
        return
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "latentsync", "models")):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (it only exists in the build container)")
    d = _mod("diffusers", _latentsync_b200_shim=True)
    d.configuration_utils = _mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin,
                                 register_to_config=register_to_config, FrozenDict=FrozenDict)
    d.models = _mod("diffusers.models", ModelMixin=ModelMixin, AutoencoderKL=object)
    d.utils = _mod("diffusers.utils", BaseOutput=BaseOutput, logging=_Logging, deprecate=lambda *a, **k: None)
    d.models.embeddings = _mod("diffusers.models.embeddings", TimestepEmbedding=TimestepEmbedding, Timesteps=Timesteps)
    d.models.attention = _mod("diffusers.models.attention", FeedForward=FeedForward, AdaLayerNorm=AdaLayerNorm)
    _mod("latentsync.utils.util", zero_rank_log=lambda logger, msg: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_unet(cfg: dict):
    """the reference's own UNet3DConditionModel built from a `model:` config dict (norm_eps cast to float)"""
    install()
    from latentsync.models.unet import UNet3DConditionModel  # noqa: the reference's class

    cfg = dict(cfg)
    cfg["norm_eps"] = float(cfg.get("norm_eps", 1e-5))
    return UNet3DConditionModel.from_config(cfg)
