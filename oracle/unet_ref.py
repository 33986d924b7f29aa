"""TEST INFRASTRUCTURE - CPU restatement ("port") of the reference UNet3DConditionModel.forward in plain PyTorch fp32.

It exists because /root/reference does not travel to the GPU box: there the CUDA path is checked against this port and
against the committed golden vectors (tests/golden/), which were produced by the reference's own modules
(oracle/make_golden.py).  The port itself is pinned against those same reference modules: make_golden.py asserts
rel-L2 < 1e-5 between the two at the full stage2 config and tests/test_oracle.py re-checks it against the fixtures.

Each function cites the reference lines it follows.  Layout is the reference's: (b, c, f, h, w) fp32.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from latentsync_b200.spec import unet_config

SD = Dict[str, torch.Tensor]

# The reference calls F.scaled_dot_product_attention (attention.py:271, motion_module.py:300).  The CPU oracle spells the
# same arithmetic out (softmax(q k^T d^-0.5) v) so that it does not depend on a fused kernel's rounding; bench.py's
# PyTorch-eager-on-GPU baseline leg sets this flag to run the reference's actual call (flash / cuDNN SDPA in fp16).
USE_SDPA = False


def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    """diffusers get_timestep_embedding as used at unet.py:95,376 (320 ch, flip_sin_to_cos, shift 0, fp32)"""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    e = t[:, None].float() * torch.exp(exponent)[None]
    emb = torch.cat([torch.sin(e), torch.cos(e)], -1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], -1)
    return emb


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv2d_bf(sd: SD, p: str, x, stride=1, padding=1):
    """InflatedConv3d.forward (resnet.py:10-18): 2-D conv over (b f)"""
    b, c, f, h, w = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w), sd[p + ".weight"], sd[p + ".bias"], stride=stride,
                 padding=padding)
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def resnet_block(sd: SD, p: str, x, emb, groups: int, eps: float):
    """ResnetBlock3D.forward (resnet.py:182-223); GroupNorm on the 5-D tensor => statistics over (C/g, F, H, W)"""
    h = F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = _conv2d_bf(sd, p + ".conv1", F.silu(h))
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(emb))[:, :, None, None, None]
    h = F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = _conv2d_bf(sd, p + ".conv2", F.silu(h))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv2d_bf(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention(sd: SD, p: str, x, ctx, heads: int):
    """Attention.forward (attention.py:250-280): q/k/v without bias, SDPA with scale d^-0.5, to_out[0] with bias"""
    ctx = x if ctx is None else ctx
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    b, s, c = q.shape
    d = c // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)

    if USE_SDPA:
        a = F.scaled_dot_product_attention(split(q), split(k), split(v))
    else:
        a = torch.softmax(split(q) @ split(k).transpose(-1, -2) * d ** -0.5, dim=-1) @ split(v)
    a = a.permute(0, 2, 1, 3).reshape(b, s, c)
    return _lin(sd, p + ".to_out.0", a)


def feed_forward(sd: SD, p: str, x):
    """diffusers FeedForward / GEGLU (attention.py:171,197): Linear(C->8C) -> value * gelu_erf(gate) -> Linear(4C->C)"""
    hg = _lin(sd, p + ".net.0.proj", x)
    h, g = hg.chunk(2, dim=-1)
    return _lin(sd, p + ".net.2", h * F.gelu(g))


def transformer3d(sd: SD, p: str, x, audio, heads: int, groups: int):
    """Transformer3DModel.forward (attention.py:82-124) + BasicTransformerBlock.forward (:174-199)"""
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = xf
    hs = F.group_norm(xf, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    hs = F.conv2d(hs, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    t = p + ".transformer_blocks.0"
    n = F.layer_norm(hs, (c,), sd[t + ".norm1.weight"], sd[t + ".norm1.bias"])
    hs = attention(sd, t + ".attn1", n, None, heads) + hs
    if (t + ".attn2.to_q.weight") in sd and audio is not None:
        ctx = audio.reshape(-1, audio.shape[-2], audio.shape[-1]) if audio.dim() == 4 else audio
        n = F.layer_norm(hs, (c,), sd[t + ".norm2.weight"], sd[t + ".norm2.bias"])
        hs = attention(sd, t + ".attn2", n, ctx, heads) + hs
    n = F.layer_norm(hs, (c,), sd[t + ".norm3.weight"], sd[t + ".norm3.bias"])
    hs = feed_forward(sd, t + ".ff", n) + hs
    hs = hs.reshape(b * f, h, w, c).permute(0, 3, 1, 2)
    hs = F.conv2d(hs, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + res
    return hs.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def motion_module(sd: SD, p: str, x, heads: int, groups: int):
    """VanillaTemporalModule -> TemporalTransformer3DModel.forward (motion_module.py:126-151) ->
    TemporalTransformerBlock.forward (:203-218) -> VersatileAttention.forward (:262-313)"""
    b, c, f, h, w = x.shape
    t = p + ".temporal_transformer"
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = xf
    hs = F.group_norm(xf, groups, sd[t + ".norm.weight"], sd[t + ".norm.bias"], 1e-6)
    hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    hs = _lin(sd, t + ".proj_in", hs)
    i = 0
    while f"{t}.transformer_blocks.{i}.ff_norm.weight" in sd:
        blk = f"{t}.transformer_blocks.{i}"
        k = 0
        while f"{blk}.attention_blocks.{k}.to_q.weight" in sd:
            a = f"{blk}.attention_blocks.{k}"
            n = F.layer_norm(hs, (c,), sd[f"{blk}.norms.{k}.weight"], sd[f"{blk}.norms.{k}.bias"])
            # "(b f) s c -> (b s) f c", + PE on the normed tokens (feeds q, k and v), attention over frames
            n = n.reshape(b, f, h * w, c).permute(0, 2, 1, 3).reshape(b * h * w, f, c)
            if (a + ".pos_encoder.pe") in sd:
                n = n + sd[a + ".pos_encoder.pe"][:, :f]
            o = attention(sd, a, n, None, heads)
            o = o.reshape(b, h * w, f, c).permute(0, 2, 1, 3).reshape(b * f, h * w, c)
            hs = o + hs
            k += 1
        n = F.layer_norm(hs, (c,), sd[blk + ".ff_norm.weight"], sd[blk + ".ff_norm.bias"])
        hs = feed_forward(sd, blk + ".ff", n) + hs
        i += 1
    hs = _lin(sd, t + ".proj_out", hs)
    hs = hs.reshape(b * f, h, w, c).permute(0, 3, 1, 2) + res
    return hs.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def upsample(sd: SD, p: str, x):
    """Upsample3D.forward (resnet.py:47-75): nearest x2 on (h, w) then 3x3 conv"""
    x = F.interpolate(x, scale_factor=[1.0, 2.0, 2.0], mode="nearest")
    return _conv2d_bf(sd, p + ".conv", x)


@torch.no_grad()
def unet_forward(sd: SD, cfg: dict, sample: torch.Tensor, timestep, audio: torch.Tensor = None, taps: dict = None):
    """UNet3DConditionModel.forward (unet.py:312-471). sample (B, Cin, F, H, W); audio (B, F, S, D) or (B*F, S, D).
    `taps`, if given, receives named intermediate activations (Appendix-B tape order) for per-block parity."""
    c = unet_config(cfg)
    g, eps = c["norm_num_groups"], c["norm_eps"]
    heads = c["attention_head_dim"]  # passed as the NUMBER of heads (unet_blocks.py:207-208)
    mm_heads = c["motion_module_kwargs"].get("num_attention_heads", 8)
    boc = c["block_out_channels"]
    nlev = len(boc)
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
    t = timestep.reshape(-1).float().expand(sample.shape[0])
    # fp32 sinusoid, then the model's dtype (unet.py:376-381: `t_emb = t_emb.to(dtype=self.dtype)`); a no-op for the fp32
    # CPU oracle, needed when bench.py runs this port in fp16 on the GPU as the PyTorch-eager baseline
    emb = timestep_embedding(t, boc[0], c["flip_sin_to_cos"], c["freq_shift"]).to(sample.dtype)
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", emb)))

    def tap(name, v):
        if taps is not None:
            taps[name] = v

    x = _conv2d_bf(sd, "conv_in", sample)
    tap("conv_in", x)
    skips = [x]
    for i, typ in enumerate(c["down_block_types"]):
        p = f"down_blocks.{i}"
        for j in range(c["layers_per_block"]):
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb, g, eps)
            if typ == "CrossAttnDownBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, audio, heads, g)
            if f"{p}.motion_modules.{j}.temporal_transformer.norm.weight" in sd:
                x = motion_module(sd, f"{p}.motion_modules.{j}", x, mm_heads, g)
            tap(f"{p}.{j}", x)
            skips.append(x)
        if i != nlev - 1:
            x = _conv2d_bf(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=c["downsample_padding"])
            skips.append(x)
    # mid (unet_blocks.py:247-260)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb, g, eps)
    x = transformer3d(sd, "mid_block.attentions.0", x, audio, heads, g)
    if "mid_block.motion_modules.0.temporal_transformer.norm.weight" in sd:
        x = motion_module(sd, "mid_block.motion_modules.0", x, mm_heads, g)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb, g, eps)
    tap("mid", x)
    for i, typ in enumerate(c["up_block_types"]):
        p = f"up_blocks.{i}"
        for j in range(c["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)  # unet_blocks.py:624,745
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb, g, eps)
            if typ == "CrossAttnUpBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, audio, heads, g)
            if f"{p}.motion_modules.{j}.temporal_transformer.norm.weight" in sd:
                x = motion_module(sd, f"{p}.motion_modules.{j}", x, mm_heads, g)
            tap(f"{p}.{j}", x)
        if i != nlev - 1:
            x = upsample(sd, f"{p}.upsamplers.0", x)
    x = F.group_norm(x, g, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps)
    return _conv2d_bf(sd, "conv_out", F.silu(x))
