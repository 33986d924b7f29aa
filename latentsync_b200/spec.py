"""Parameter inventory (state_dict key -> shape) of the two networks on the hot path.

The key names are the reference's on-disk checkpoint contract: they are what
`UNet3DConditionModel.load_state_dict` (latentsync/models/unet.py:473-492) and diffusers' AutoencoderKL accept.
The UNet list is derived from the constructor logic in unet.py:85-241 / unet_blocks.py / resnet.py:104-180 /
attention.py:23-80,127-172,202-235 / motion_module.py:39-124,154-201,221-260; tests check it against the state_dict
of the reference's own module (1 246 entries for configs/unet/stage2.yaml).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

Shape = Tuple[int, ...]

# configs/unet/stage2.yaml `model:` section (the LatentSync 1.5 inference config, inference.sh:3-10)
STAGE2_UNET_CONFIG = dict(
    act_fn="silu",
    add_audio_layer=True,
    attention_head_dim=8,
    block_out_channels=(320, 640, 1280, 1280),
    center_input_sample=False,
    cross_attention_dim=384,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    downsample_padding=1,
    flip_sin_to_cos=True,
    freq_shift=0,
    in_channels=13,
    layers_per_block=2,
    mid_block_scale_factor=1,
    norm_eps=1e-5,
    norm_num_groups=32,
    out_channels=4,
    sample_size=64,
    resnet_time_scale_shift="default",
    use_motion_module=True,
    motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=False,
    motion_module_decoder_only=False,
    motion_module_type="Vanilla",
    motion_module_kwargs=dict(
        num_attention_heads=8,
        num_transformer_block=1,
        attention_block_types=("Temporal_Self", "Temporal_Self"),
        temporal_position_encoding=True,
        temporal_position_encoding_max_len=24,
        temporal_attention_dim_div=1,
        zero_initialize=True,
    ),
)

# same topology at a quarter of the width: fast enough for CPU-side oracle runs inside the test-suite
TINY_UNET_CONFIG = dict(STAGE2_UNET_CONFIG, block_out_channels=(128, 256, 256, 256))

# constructor defaults of UNet3DConditionModel.__init__ (unet.py:43-83)
UNET_CTOR_DEFAULTS = dict(
    sample_size=None,
    in_channels=4,
    out_channels=4,
    center_input_sample=False,
    flip_sin_to_cos=True,
    freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    only_cross_attention=False,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    downsample_padding=1,
    mid_block_scale_factor=1,
    act_fn="silu",
    norm_num_groups=32,
    norm_eps=1e-5,
    cross_attention_dim=1280,
    attention_head_dim=8,
    dual_cross_attention=False,
    use_linear_projection=False,
    class_embed_type=None,
    num_class_embeds=None,
    upcast_attention=False,
    resnet_time_scale_shift="default",
    use_inflated_groupnorm=False,
    use_motion_module=False,
    motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=False,
    motion_module_decoder_only=False,
    motion_module_type=None,
    motion_module_kwargs={},
    add_audio_layer=False,
)

# stabilityai/sd-vae-ft-mse AutoencoderKL config (scripts/inference.py:56-58 overrides scaling/shift)
SD_VAE_FT_MSE_CONFIG = dict(
    in_channels=3,
    out_channels=3,
    latent_channels=4,
    block_out_channels=(128, 256, 512, 512),
    layers_per_block=2,
    norm_num_groups=32,
    act_fn="silu",
    sample_size=256,
    scaling_factor=0.18215,
    shift_factor=0.0,
)


def _norm(d: Dict[str, Shape], p: str, c: int) -> None:
    d[p + ".weight"] = (c,)
    d[p + ".bias"] = (c,)


def _conv(d, p, cout, cin, k) -> None:
    d[p + ".weight"] = (cout, cin, k, k)
    d[p + ".bias"] = (cout,)


def _linear(d, p, cout, cin, bias=True) -> None:
    d[p + ".weight"] = (cout, cin)
    if bias:
        d[p + ".bias"] = (cout,)


def _resnet3d(d, p, cin, cout, temb) -> None:
    """ResnetBlock3D (resnet.py:104-180)"""
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3)
    _linear(d, p + ".time_emb_proj", cout, temb)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, p + ".conv_shortcut", cout, cin, 1)


def _feed_forward(d, p, c) -> None:
    """diffusers FeedForward(dim, activation_fn='geglu'): net.0 = GEGLU(proj: C -> 8C), net.2 = Linear(4C -> C)"""
    _linear(d, p + ".net.0.proj", 8 * c, c)
    _linear(d, p + ".net.2", c, 4 * c)


def _attention(d, p, c, kv_dim) -> None:
    """Attention (attention.py:202-235): to_q/k/v without bias, to_out.0 with bias"""
    _linear(d, p + ".to_q", c, c, bias=False)
    _linear(d, p + ".to_k", c, kv_dim, bias=False)
    _linear(d, p + ".to_v", c, kv_dim, bias=False)
    _linear(d, p + ".to_out.0", c, c)


def _transformer3d(d, p, c, cross_dim, audio) -> None:
    """Transformer3DModel with one BasicTransformerBlock (attention.py:23-80,127-172)"""
    _norm(d, p + ".norm", c)
    _conv(d, p + ".proj_in", c, c, 1)
    b = p + ".transformer_blocks.0"
    _norm(d, b + ".norm1", c)
    _attention(d, b + ".attn1", c, c)
    if audio:
        _norm(d, b + ".norm2", c)
        _attention(d, b + ".attn2", c, cross_dim)
    _feed_forward(d, b + ".ff", c)
    _norm(d, b + ".norm3", c)
    _conv(d, p + ".proj_out", c, c, 1)


def _motion(d, p, c, kw) -> None:
    """VanillaTemporalModule (motion_module.py:39-124,154-201,237-260)"""
    t = p + ".temporal_transformer"
    _norm(d, t + ".norm", c)
    _linear(d, t + ".proj_in", c, c)
    nblk = kw.get("num_transformer_block", 2)
    types = kw.get("attention_block_types", ("Temporal_Self", "Temporal_Self"))
    max_len = kw.get("temporal_position_encoding_max_len", 24)
    for i in range(nblk):
        b = f"{t}.transformer_blocks.{i}"
        for k in range(len(types)):
            _attention(d, f"{b}.attention_blocks.{k}", c, c)
            if kw.get("temporal_position_encoding", False):
                d[f"{b}.attention_blocks.{k}.pos_encoder.pe"] = (1, max_len, c)
        for k in range(len(types)):
            _norm(d, f"{b}.norms.{k}", c)
        _feed_forward(d, b + ".ff", c)
        _norm(d, b + ".ff_norm", c)
    _linear(d, t + ".proj_out", c, c)


def unet_config(cfg: dict) -> dict:
    """constructor kwargs merged over the reference's defaults (what diffusers' register_to_config records)"""
    out = dict(UNET_CTOR_DEFAULTS)
    for k, v in cfg.items():
        if k in out:
            out[k] = v
    out["norm_eps"] = float(out["norm_eps"])  # yaml.safe_load reads `1e-5` as a string
    for k in ("block_out_channels", "down_block_types", "up_block_types", "motion_module_resolutions"):
        out[k] = tuple(out[k])
    return out


# Constructor options that the kernel plan does not implement: they must sit at the value below (all of them do in
# configs/unet/stage2.yaml and stage1.yaml); anything else raises instead of silently computing something different.
_SUPPORTED_ONLY = dict(
    only_cross_attention=False,      # unet_blocks.py passes it to Transformer3DModel: attn1 would become cross-attention
    dual_cross_attention=False,
    upcast_attention=False,          # attention runs with fp32 scores / softmax anyway, but to_q / to_k stay fp16
    use_linear_projection=False,
    use_inflated_groupnorm=False,    # True = per-frame GroupNorm statistics in the resnets (resnet.py:21-30)
    mid_block_scale_factor=1,        # output_scale_factor of the mid-block resnets (unet_blocks.py:176)
    act_fn="silu",
    downsample_padding=1,
    resnet_time_scale_shift="default",
    class_embed_type=None,
    num_class_embeds=None,
    mid_block_type="UNetMidBlock3DCrossAttn",
)


def validate_unet_config(cfg: dict) -> None:
    """raise NotImplementedError for every constructor option the engine would otherwise ignore"""
    c = unet_config(cfg)
    for k, want in _SUPPORTED_ONLY.items():
        if c[k] != want:
            raise NotImplementedError(f"UNet3DConditionModel({k}={c[k]!r}) is not implemented by latentsync_b200 "
                                      f"(only {k}={want!r}: the LatentSync stage1/stage2 configs)")
    if c["use_motion_module"] and c["motion_module_type"] not in ("Vanilla", None):
        raise NotImplementedError(f"motion_module_type={c['motion_module_type']!r} (only 'Vanilla')")
    kw = c["motion_module_kwargs"] or {}
    if c["use_motion_module"]:
        types = tuple(kw.get("attention_block_types", ("Temporal_Self", "Temporal_Self")))
        if any(t != "Temporal_Self" for t in types):
            raise NotImplementedError(f"attention_block_types={types!r} (only 'Temporal_Self')")
        if kw.get("temporal_attention_dim_div", 1) != 1:
            raise NotImplementedError("temporal_attention_dim_div != 1")
    for t in c["down_block_types"]:
        if t not in ("CrossAttnDownBlock3D", "DownBlock3D"):
            raise NotImplementedError(f"down block type {t!r}")
    for t in c["up_block_types"]:
        if t not in ("CrossAttnUpBlock3D", "UpBlock3D"):
            raise NotImplementedError(f"up block type {t!r}")


def unet_param_spec(cfg: dict) -> "OrderedDict[str, Shape]":
    """state_dict keys and shapes of UNet3DConditionModel(**cfg), in the reference's registration order."""
    c = unet_config(cfg)
    assert c["mid_block_type"] == "UNetMidBlock3DCrossAttn"
    assert c["resnet_time_scale_shift"] == "default" and not c["use_linear_projection"]
    assert c["class_embed_type"] is None and c["num_class_embeds"] is None
    boc = c["block_out_channels"]
    temb = boc[0] * 4
    audio = c["add_audio_layer"]
    cross = c["cross_attention_dim"]
    mm = c["use_motion_module"]
    mm_res = c["motion_module_resolutions"]
    kw = c["motion_module_kwargs"]
    d: "OrderedDict[str, Shape]" = OrderedDict()
    _conv(d, "conv_in", boc[0], c["in_channels"], 3)
    _linear(d, "time_embedding.linear_1", temb, boc[0])
    _linear(d, "time_embedding.linear_2", temb, temb)
    nlev = len(boc)
    out_ch = boc[0]
    for i, typ in enumerate(c["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        p = f"down_blocks.{i}"
        has_attn = typ == "CrossAttnDownBlock3D"
        has_mm = mm and (2 ** i in mm_res) and not c["motion_module_decoder_only"]
        if has_attn:
            for j in range(c["layers_per_block"]):
                _transformer3d(d, f"{p}.attentions.{j}", out_ch, cross, audio)
        for j in range(c["layers_per_block"]):
            _resnet3d(d, f"{p}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch, temb)
        if has_mm:
            for j in range(c["layers_per_block"]):
                _motion(d, f"{p}.motion_modules.{j}", out_ch, kw)
        if i != nlev - 1:
            _conv(d, f"{p}.downsamplers.0.conv", out_ch, out_ch, 3)
    rev = list(reversed(boc))
    out_ch = rev[0]
    for i, typ in enumerate(c["up_block_types"]):
        prev, out_ch = out_ch, rev[i]
        in_ch = rev[min(i + 1, nlev - 1)]
        p = f"up_blocks.{i}"
        has_attn = typ == "CrossAttnUpBlock3D"
        has_mm = mm and (2 ** (3 - i) in mm_res)
        n = c["layers_per_block"] + 1
        if has_attn:
            for j in range(n):
                _transformer3d(d, f"{p}.attentions.{j}", out_ch, cross, audio)
        for j in range(n):
            skip = in_ch if j == n - 1 else out_ch
            rin = prev if j == 0 else out_ch
            _resnet3d(d, f"{p}.resnets.{j}", rin + skip, out_ch, temb)
        if has_mm:
            for j in range(n):
                _motion(d, f"{p}.motion_modules.{j}", out_ch, kw)
        if i != nlev - 1:
            _conv(d, f"{p}.upsamplers.0.conv", out_ch, out_ch, 3)
    # mid (unet_blocks.py:153-245); registered after up_blocks because unet.py:110 pre-assigns `self.mid_block = None`
    cm = boc[-1]
    _transformer3d(d, "mid_block.attentions.0", cm, cross, audio)
    if mm and c["motion_module_mid_block"]:
        _motion(d, "mid_block.motion_modules.0", cm, kw)
    _resnet3d(d, "mid_block.resnets.0", cm, cm, temb)
    _resnet3d(d, "mid_block.resnets.1", cm, cm, temb)
    _norm(d, "conv_norm_out", boc[0])
    _conv(d, "conv_out", c["out_channels"], boc[0], 3)
    return d


def _resnet2d(d, p, cin, cout) -> None:
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, p + ".conv_shortcut", cout, cin, 1)


def vae_decoder_param_spec(cfg: dict = SD_VAE_FT_MSE_CONFIG) -> "OrderedDict[str, Shape]":
    """Decoder half of diffusers' AutoencoderKL (post_quant_conv + decoder.*), sd-vae-ft-mse layout."""
    boc = tuple(cfg["block_out_channels"])
    lat = cfg["latent_channels"]
    d: "OrderedDict[str, Shape]" = OrderedDict()
    top = boc[-1]
    _conv(d, "decoder.conv_in", top, lat, 3)
    rev = list(reversed(boc))
    out_ch = rev[0]
    for i in range(len(boc)):
        prev, out_ch = out_ch, rev[i]
        for j in range(cfg["layers_per_block"] + 1):
            _resnet2d(d, f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out_ch, out_ch)
        if i != len(boc) - 1:
            _conv(d, f"decoder.up_blocks.{i}.upsamplers.0.conv", out_ch, out_ch, 3)
    a = "decoder.mid_block.attentions.0"
    _norm(d, a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _linear(d, f"{a}.{n}", top, top)
    _resnet2d(d, "decoder.mid_block.resnets.0", top, top)
    _resnet2d(d, "decoder.mid_block.resnets.1", top, top)
    _norm(d, "decoder.conv_norm_out", boc[0])
    _conv(d, "decoder.conv_out", cfg["out_channels"], boc[0], 3)
    _conv(d, "post_quant_conv", lat, lat, 1)
    return d


def vae_encoder_param_spec(cfg: dict = SD_VAE_FT_MSE_CONFIG) -> "OrderedDict[str, Shape]":
    """Encoder half of diffusers' AutoencoderKL (encoder.* + quant_conv), sd-vae-ft-mse layout: conv_in,
    DownEncoderBlock2D x4 (2 ResnetBlock2D each, stride-2 conv after the first three), UNetMidBlock2D,
    GroupNorm-SiLU-conv_out to 2*latent channels, 1x1 quant_conv.  Call sites: lipsync_pipeline.py:298,315."""
    boc = tuple(cfg["block_out_channels"])
    lat = cfg["latent_channels"]
    d: "OrderedDict[str, Shape]" = OrderedDict()
    _conv(d, "encoder.conv_in", boc[0], cfg.get("in_channels", 3), 3)
    out_ch = boc[0]
    for i in range(len(boc)):
        prev, out_ch = out_ch, boc[i]
        for j in range(cfg["layers_per_block"]):
            _resnet2d(d, f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else out_ch, out_ch)
        if i != len(boc) - 1:
            _conv(d, f"encoder.down_blocks.{i}.downsamplers.0.conv", out_ch, out_ch, 3)
    top = boc[-1]
    a = "encoder.mid_block.attentions.0"
    _norm(d, a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _linear(d, f"{a}.{n}", top, top)
    _resnet2d(d, "encoder.mid_block.resnets.0", top, top)
    _resnet2d(d, "encoder.mid_block.resnets.1", top, top)
    _norm(d, "encoder.conv_norm_out", top)
    _conv(d, "encoder.conv_out", 2 * lat, top, 3)
    _conv(d, "quant_conv", 2 * lat, 2 * lat, 1)
    return d

