"""Per-kernel numerics: every C-ABI op against a plain PyTorch fp32 statement of the reference op it replaces.

Tolerances: operands are fp16 (inputs are generated in fp16 so both sides see identical values); accumulation is
fp32 on both sides, so the only difference is summation order and the final fp16 store: rel-L2 <= 2e-3.
"""
import math

import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _ops():
    from latentsync_b200 import _lib

    return _lib


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cl(x):
    """(N, C, H, W) -> channels-last rows [(N H W), C]"""
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def uncl(rows, n, h, w):
    return rows.reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


def pack_conv_w(w):
    """OIHW -> [N][(tap, c)] fp16"""
    n, c, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(n, kh * kw * c).contiguous()


# ------------------------------------------------------------------------------------------------- GEMM (Linear)
@pytest.mark.parametrize(
    "M,K,N,tile_n",
    [
        (256, 64, 32, 0),
        (128, 128, 128, 128),
        (1000, 320, 320, 0),
        (32768, 320, 960, 0),
        (2048, 1280, 1280, 256),
        (512, 1280, 1280, 160),
        (4096, 640, 2560, 64),
        (300, 384, 40, 0),
    ],
)
def test_gemm_linear(M, K, N, tile_n):
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(M, K, generator=g).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).half().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).half().to(DEV)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out, N, bias=bias, residual=res, ldr=N, tile_n=tile_n)
    ref = a.float() @ w.float().t() + bias + res.float()
    assert rel_l2(out, ref) < 2e-3
    # fp32 output, no bias/residual
    out32 = torch.empty(M, N, dtype=torch.float32, device=DEV)
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out32, N, flags=L.EPI_OUT_F32, tile_n=tile_n)
    ref = a.float() @ w.float().t()
    assert rel_l2(out32, ref) < 1e-4


def test_gemm_bias_div_and_silu():
    """per-batch-element bias rows (time embedding add, resnet.py:205) and SiLU epilogue"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(2)
    M, K, N = 2 * 1024, 128, 192
    a = torch.randn(M, K, generator=g).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).half().to(DEV)
    bias = torch.randn(2, N, generator=g).to(DEV)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out, N, bias=bias, bias_div=1024, flags=L.EPI_SILU)
    ref = a.float() @ w.float().t() + bias.repeat_interleave(1024, 0)
    ref = F.silu(ref)
    assert rel_l2(out, ref) < 2e-3


def test_gemm_geglu():
    """diffusers GEGLU: [h | g] = Linear(C -> 8C)(x); y = h * gelu_erf(g)"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(3)
    M, C = 1024, 320
    inner = 4 * C
    for tile_n in (64, 128, 256):
        a = torch.randn(M, C, generator=g).half().to(DEV)
        w = (torch.randn(2 * inner, C, generator=g) / math.sqrt(C)).half().to(DEV)
        b = torch.randn(2 * inner, generator=g).to(DEV)
        wp, bp = L.pack_geglu(w, b, tile_n)
        out = torch.empty(M, inner, dtype=torch.float16, device=DEV)
        L.gemm([L.Seg(a, C, C, 1)], 1, 1, M, wp, 2 * inner, out, inner, bias=bp, flags=L.EPI_GEGLU, tile_n=tile_n)
        hg = a.float() @ w.float().t() + b
        ref = hg[:, :inner] * F.gelu(hg[:, inner:])
        assert rel_l2(out, ref) < 2e-3, tile_n
        # with a residual (added after the gate): several N tiles, ragged M (the Whisper front end's second convolution:
        # GELU through the GEGLU epilogue + positional embedding, whisper/model.py:150-154)
        res = torch.randn(M, inner, generator=g).half().to(DEV)
        for rows in (M, M - 100):
            out2 = torch.zeros(M, inner, dtype=torch.float16, device=DEV)
            L.gemm([L.Seg(a, C, C, 1)], 1, 1, rows, wp, 2 * inner, out2, inner, bias=bp, residual=res, ldr=inner,
                   flags=L.EPI_GEGLU, tile_n=tile_n)
            assert rel_l2(out2[:rows], ref[:rows] + res[:rows].float()) < 2e-3, tile_n
            assert not out2[rows:].any()


def test_gemm_batched():
    """batched GEMM (VAE mid attention QK^T and PV): out[b] = A[b] @ B[b]^T"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(4)
    nb, M, K, N = 4, 1024, 512, 1024
    a = torch.randn(nb, M, K, generator=g).half().to(DEV)
    b = (torch.randn(nb, N, K, generator=g) / math.sqrt(K)).half().to(DEV)
    out = torch.empty(nb, M, N, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(a, K, K, 1)], nb, 1, M, b, N, out, N, b_batch_stride=N * K)
    ref = torch.bmm(a.float(), b.float().transpose(1, 2))
    assert rel_l2(out, ref) < 2e-3


# -------------------------------------------------------------------------------------------- implicit-GEMM conv
@pytest.mark.parametrize(
    "nimg,H,W,Cin,Cout",
    [
        (2, 32, 32, 64, 64),
        (32, 32, 32, 320, 320),
        (32, 16, 16, 640, 640),
        (32, 8, 8, 1280, 1280),
        (32, 4, 4, 1280, 1280),
        (3, 4, 4, 128, 96),
        (2, 64, 64, 128, 128),
        (1, 256, 256, 128, 3),
        (2, 128, 128, 64, 32),
        (5, 8, 8, 64, 40),
    ],
)
def test_conv3x3(nimg, H, W, Cin, Cout):
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(nimg, Cin, H, W, generator=g).half().to(DEV)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)).half().to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    xcl = cl(x)
    out = torch.empty(nimg * H * W, Cout, dtype=torch.float32, device=DEV)
    L.gemm([L.Seg(xcl, Cin, Cin, 9)], nimg, H, W, pack_conv_w(w), Cout, out, Cout, bias=b, flags=L.EPI_OUT_F32)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    assert rel_l2(uncl(out, nimg, H, W), ref) < 1e-4


def test_conv3x3_concat_and_shortcut():
    """skip-concat conv (unet_blocks.py:624) as two K segments + fused 1x1 shortcut (resnet.py:219-221) as a third"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(6)
    nimg, H, W, c1, c2, cs, Cout = 4, 16, 16, 128, 64, 192, 128
    x1 = torch.randn(nimg, c1, H, W, generator=g).half().to(DEV)
    x2 = torch.randn(nimg, c2, H, W, generator=g).half().to(DEV)
    xs = torch.randn(nimg, cs, H, W, generator=g).half().to(DEV)
    w = (torch.randn(Cout, c1 + c2, 3, 3, generator=g) / math.sqrt(9 * (c1 + c2))).half().to(DEV)
    ws = (torch.randn(Cout, cs, 1, 1, generator=g) / math.sqrt(cs)).half().to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    wp = torch.cat([pack_conv_w(w[:, :c1]), pack_conv_w(w[:, c1:]), ws.reshape(Cout, cs)], dim=1).contiguous()
    out = torch.empty(nimg * H * W, Cout, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(cl(x1), c1, c1, 9), L.Seg(cl(x2), c2, c2, 9), L.Seg(cl(xs), cs, cs, 1)], nimg, H, W, wp, Cout, out,
           Cout, bias=b)
    ref = F.conv2d(torch.cat([x1, x2], 1).float(), w.float(), b, padding=1) + F.conv2d(xs.float(), ws.float())
    assert rel_l2(uncl(out, nimg, H, W), ref) < 2e-3


@pytest.mark.parametrize("nimg,H,W,C,N", [(32, 8, 8, 1280, 1280), (32, 16, 16, 640, 640), (4, 32, 32, 512, 512),
                                          (2, 64, 64, 256, 256), (1, 128, 128, 128, 128), (3, 16, 16, 128, 96)])
def test_upsample_folded_into_conv(nimg, H, W, C, N):
    """Upsample3D / diffusers Upsample2D (resnet.py:47-75: F.interpolate(scale_factor=2, mode="nearest") then a 3x3
    convolution) as four sub-pixel phase GEMMs over the LOW-resolution tensor (LsGemmArgs.up2): 2 x 2 windows, weights =
    sums of the 3x3 taps that land on the same source pixel, each launch writes every second pixel / row of the output"""
    L = _ops()
    from latentsync_b200.engine import pack_upconv_phases

    g = torch.Generator(device="cpu").manual_seed(41)
    x = torch.randn(nimg, C, H, W, generator=g).half().to(DEV)
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    phases = pack_upconv_phases(w)
    out = torch.full((nimg * 4 * H * W, N), float("nan"), dtype=torch.float16, device=DEV)
    for ph in range(4):
        L.gemm([L.Seg(cl(x), C, C, 4)], nimg, H, W, phases[ph], N, out, N, bias=b, up2=ph + 1)
    ref = F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w.half().float(), b, padding=1)
    got = uncl(out, nimg, 2 * H, 2 * W)
    assert not torch.isnan(got).any()
    assert rel_l2(got, ref) < 2e-3
    with pytest.raises(RuntimeError):  # a residual cannot follow the strided store
        L.gemm([L.Seg(cl(x), C, C, 4)], nimg, H, W, phases[0], N, out, N, bias=b, residual=out, ldr=N, up2=1)


@pytest.mark.parametrize("nimg,H,W,C,N", [(32, 16, 16, 320, 320), (32, 8, 8, 640, 640), (32, 4, 4, 1280, 1280),
                                          (2, 64, 64, 128, 128), (1, 128, 128, 128, 128), (3, 32, 32, 256, 96)])
@pytest.mark.parametrize("pad", [1, 0])
def test_conv_stride2_in_place(nimg, H, W, C, N, pad):
    """Downsample3D (resnet.py:78-101: Conv2d 3x3, stride 2, padding 1) and diffusers Downsample2D of the VAE encoder
    (F.pad(x, (0, 1, 0, 1)) + stride-2 conv) WITHOUT an im2col copy: LsGemmArgs.stride2 - output geometry, the 9-tap
    segment points at the [nimg, 2H, 2W] input, TMA element strides of two fetch every second pixel / row"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(43)
    x = torch.randn(nimg, C, 2 * H, 2 * W, generator=g).half().to(DEV)
    w = (torch.randn(N, C, 3, 3, generator=g) / math.sqrt(9 * C)).half().to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    out = torch.full((nimg * H * W, N), float("nan"), dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(cl(x), C, C, 9)], nimg, H, W, pack_conv_w(w), N, out, N, bias=b, stride2=1, stride2_pad=pad)
    xf = x.float() if pad else F.pad(x.float(), (0, 1, 0, 1))
    ref = F.conv2d(xf, w.float(), b, stride=2, padding=pad)
    assert ref.shape[-2:] == (H, W)
    assert rel_l2(uncl(out, nimg, H, W), ref) < 2e-3


def test_conv_stride2_via_im2col():
    """Downsample3D (resnet.py:89): 3x3 stride 2 pad 1"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(7)
    nimg, H, W, C, Cout = 4, 16, 16, 64, 64
    x = torch.randn(nimg, C, H, W, generator=g).half().to(DEV)
    w = (torch.randn(Cout, C, 3, 3, generator=g) / math.sqrt(9 * C)).half().to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    cols = torch.empty(nimg * (H // 2) * (W // 2), 9 * C, dtype=torch.float16, device=DEV)
    L.im2col_s2(cl(x), nimg, H, W, C, cols)
    out = torch.empty(nimg * (H // 2) * (W // 2), Cout, dtype=torch.float16, device=DEV)
    M = cols.shape[0]
    L.gemm([L.Seg(cols, 9 * C, 9 * C, 1)], 1, 1, M, pack_conv_w(w), Cout, out, Cout, bias=b)
    ref = F.conv2d(x.float(), w.float(), b, stride=2, padding=1)
    assert rel_l2(uncl(out, nimg, H // 2, W // 2), ref) < 2e-3


def test_upsample2x():
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(8)
    x = torch.randn(3, 64, 8, 4, generator=g).half().to(DEV)
    y = torch.empty(3 * 16 * 8, 64, dtype=torch.float16, device=DEV)
    L.upsample2x(cl(x), 3, 8, 4, 64, y)
    ref = F.interpolate(x.float(), scale_factor=2.0, mode="nearest")
    assert torch.equal(uncl(y, 3, 16, 8).float(), ref)


# --------------------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("c1,c2,rows_per_inst,ninst", [(320, 0, 16 * 1024, 2), (1280, 640, 64, 32), (640, 320, 256, 4),
                                                       (128, 0, 65536, 2), (512, 0, 1024, 3),
                                                       # thread-block-cluster path (instance fits <= 16 CTAs' smem):
                                                       # per-frame norms of each level, 4x4 concat, 8x8 joint (16 CTAs)
                                                       (320, 0, 1024, 32), (640, 0, 256, 32), (1280, 0, 16, 32),
                                                       (1280, 1280, 256, 2), (1280, 0, 1024, 2), (1280, 0, 8, 3)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(c1, c2, rows_per_inst, ninst, silu):
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(9)
    C = c1 + c2
    rows = rows_per_inst * ninst
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).half().to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    x1 = x[:, :c1].contiguous()
    x2 = x[:, c1:].contiguous() if c2 else None
    out = torch.empty(rows, C, dtype=torch.float16, device=DEV)
    stats = torch.empty(ninst * 32 * 2, dtype=torch.float32, device=DEV)
    L.groupnorm(x1, c1, x2, c2, rows, rows_per_inst, 32, gamma, beta, 1e-5, silu, out, stats)
    xr = x.float().reshape(ninst, rows_per_inst, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(rows, C)
    assert rel_l2(out, ref) < 2e-3
    out2 = torch.empty_like(out)
    L.groupnorm_fused(x1, c1, x2, c2, rows, rows_per_inst, 32, gamma, beta, 1e-5, silu, out2, stats)
    assert rel_l2(out2, ref) < 2e-3
    out3 = torch.empty_like(out)
    L.groupnorm_fused(x1, c1, x2, c2, rows, rows_per_inst, 32, gamma, beta, 1e-5, silu, out3, stats)
    assert torch.equal(out2, out3)  # deterministic, and the self-resetting tickets survive a relaunch


@pytest.mark.parametrize("M,K,N,unit,pair,conv", [
    (32768, 320, 320, 10, 1, False),    # level-0 short-K linear: >= 3 tiles per CTA (alternate tiles), 160-wide tiles
    (32768, 320, 320, 10, 2, False),    # CTA pairs
    (2048, 1280, 1280, 10, 0, False),   # both epilogue groups share every tile (3 + 2 slabs, trailing single slab)
    (8192, 640, 640, 10, 0, True),      # 3x3 convolution, 16 x 16 images
    (4096, 128, 256, 8, 0, False),      # VAE-like: unit 8, any tile width dividing N
    (16384, 256, 128, 4, 0, True),      # 128-wide tiles, unit 4
])
def test_groupnorm_partials_from_gemm_epilogue(M, K, N, unit, pair, conv):
    """nn.GroupNorm without a statistics pass (resnet.py:185-187,207-215; attention.py:96; motion_module.py:139): the GEMM
    that produces x writes per 128-row tile and per `unit` columns the (sum, sum of squares) of the fp16 values it stores;
    ls_groupnorm_parts sums them per (instance, group) and normalises in one pass.  Checked: the partials against the
    stored tensor, the norm against F.group_norm for per-frame and joint instances, bitwise determinism, and that the
    partials do not depend on the launch geometry (the same rows produced by two half-size launches)."""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(33)
    bias = (0.5 + 0.3 * torch.randn(N, generator=g)).to(DEV)
    res = (torch.randn(M, N, generator=g) * 1.5).half().to(DEV)
    x = torch.empty(M, N, dtype=torch.float16, device=DEV)
    U = N // unit
    parts = torch.full((M // 128, U, 2), float("nan"), dtype=torch.float32, device=DEV)
    if conv:
        side = 16
        nimg = M // (side * side)
        a = torch.randn(nimg, K, side, side, generator=g).half().to(DEV)
        w = (torch.randn(N, K, 3, 3, generator=g) / (9 * K) ** 0.5).half().to(DEV)
        segs, geo, wp = [L.Seg(cl(a), K, K, 9)], (nimg, side, side), pack_conv_w(w)
        want = cl(F.conv2d(a.float(), w.float(), padding=1)) + bias + res.float()
    else:
        a = torch.randn(M, K, generator=g).half().to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(DEV)
        segs, geo, wp = [L.Seg(a, K, K, 1)], (1, 1, M), w
        want = a.float() @ w.float().t() + bias + res.float()
    L.gemm(segs, *geo, wp, N, x, N, bias=bias, residual=res, ldr=N, cta_pair=pair, gn_partials_out=parts, gn_unit=unit)
    assert rel_l2(x, want) < 1e-3
    assert not torch.isnan(parts).any()
    xf = x.float().reshape(M // 128, 128, U, unit)
    assert torch.allclose(parts[..., 0], xf.sum((1, 3)), atol=2e-2, rtol=1e-5)
    assert torch.allclose(parts[..., 1], (xf * xf).sum((1, 3)), rtol=1e-4)
    # same launch again: bitwise; the two halves of the rows as separate launches (other tile-to-CTA assignment): bitwise
    parts2 = torch.full_like(parts, float("nan"))
    x2 = torch.empty_like(x)
    L.gemm(segs, *geo, wp, N, x2, N, bias=bias, residual=res, ldr=N, cta_pair=pair, gn_partials_out=parts2, gn_unit=unit)
    assert torch.equal(parts, parts2) and torch.equal(x, x2)
    if not conv:
        half = M // 2
        parts3 = torch.full_like(parts, float("nan"))
        for e in range(2):
            L.gemm([L.Seg(a[e * half:], K, K, 1)], 1, 1, half, w, N, x2[e * half:], N, bias=bias, residual=res[e * half:],
                   ldr=N, gn_partials_out=parts3[e * (half // 128):], gn_unit=unit)
        assert torch.equal(parts, parts3) and torch.equal(x, x2)
    # the norm: per-frame-like instances (1024 rows) and one joint instance per half, channels-per-group = N / 32
    gamma = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(N, generator=g)).to(DEV)
    for rpi in (1024, M // 2):
        for silu in (False, True):
            out = torch.empty_like(x)
            L.groupnorm_parts(x, N, parts, None, 0, None, M, rpi, 32, unit, gamma, beta, 1e-5, silu, out)
            ref = F.group_norm(x.float().reshape(M // rpi, rpi, N).permute(0, 2, 1), 32, gamma, beta, 1e-5)
            ref = (F.silu(ref) if silu else ref).permute(0, 2, 1).reshape(M, N)
            assert rel_l2(out, ref) < 2e-3
            out_b = torch.empty_like(x)
            L.groupnorm_parts(x, N, parts, None, 0, None, M, rpi, 32, unit, gamma, beta, 1e-5, silu, out_b)
            assert torch.equal(out, out_b)
    with pytest.raises(RuntimeError):  # an explicit tile width must hold whole units and divide N
        L.gemm(segs, *geo, wp, N, x2, N, bias=bias, tile_n=96, gn_partials_out=parts2, gn_unit=unit)


@pytest.mark.parametrize("c1,c2,rpi,ninst", [(640, 320, 16384, 2), (1280, 640, 1024, 2), (320, 320, 256, 8),
                                             (1280, 1280, 1024, 2)])
def test_groupnorm_parts_of_a_virtual_concat(c1, c2, rpi, ninst):
    """norm1 of the up blocks reads torch.cat([hidden, skip]) (unet_blocks.py:624,745): groups straddle the two sources
    (960 channels: 30 per group), each source brings the partials of its own producer; wide tensors are split over the
    channels as well as the rows (csplit)"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(35)
    rows, unit = rpi * ninst, 10
    srcs = []
    for c in (c1, c2):
        a = torch.randn(rows, 64, generator=g).half().to(DEV)
        w = (torch.randn(c, 64, generator=g) / 8).half().to(DEV)
        b = (0.3 * torch.randn(c, generator=g)).to(DEV)
        x = torch.empty(rows, c, dtype=torch.float16, device=DEV)
        parts = torch.empty(rows // 128, c // unit, 2, dtype=torch.float32, device=DEV)
        L.gemm([L.Seg(a, 64, 64, 1)], 1, 1, rows, w, c, x, c, bias=b, gn_partials_out=parts, gn_unit=unit)
        srcs.append((x, parts))
    C = c1 + c2
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    out = torch.empty(rows, C, dtype=torch.float16, device=DEV)
    L.groupnorm_parts(srcs[0][0], c1, srcs[0][1], srcs[1][0], c2, srcs[1][1], rows, rpi, 32, unit, gamma, beta, 1e-5, True, out)
    xx = torch.cat([srcs[0][0], srcs[1][0]], 1).float()
    ref = F.silu(F.group_norm(xx.reshape(ninst, rpi, C).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel_l2(out, ref.reshape(rows, C)) < 2e-3


@pytest.mark.parametrize("C", [320, 640, 1280])
def test_layernorm_and_pe(C):
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(10)
    Fr, HW, B = 16, 20, 2
    rows = B * Fr * HW
    x = (torch.randn(rows, C, generator=g) * 3 + 1).half().to(DEV)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    out = torch.empty_like(x)
    L.layernorm(x, rows, C, gamma, beta, 1e-5, out)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert rel_l2(out, ref) < 1e-3
    pe = torch.randn(Fr, C, generator=g).to(DEV)
    L.layernorm(x, rows, C, gamma, beta, 1e-5, out, pe=pe, rows_per_frame=HW, nframes=Fr)
    ref2 = (ref.reshape(B, Fr, HW, C) + pe[None, :, None, :]).reshape(rows, C)
    assert rel_l2(out, ref2) < 1e-3


@pytest.mark.parametrize("M,C,N,mode,pair", [(33000, 320, 960, "plain", 1), (33000, 320, 960, "plain", 2),
                                             (1000, 640, 640, "plain", 0), (2048, 1280, 3840, "plain", 0),
                                             (33000, 320, 2560, "geglu", 0), (2048, 1280, 10240, "geglu", 2),
                                             (32768, 320, 960, "pe", 0), (8192, 640, 1920, "pe", 2)])
def test_layernorm_folded_into_gemm(M, C, N, mode, pair):
    """nn.LayerNorm -> nn.Linear (attention.py:176-186,196-197; motion_module.py:208-216,232-234) with NO LayerNorm pass:
    the GEMM that produces x (bias + residual epilogue) emits per-row (sum, sum of squares) partials of the fp16 values
    it stores, the consuming GEMM runs on the raw x with W diag(gamma) and applies
    out = rstd (x W'^T) - rstd mean col_sum + (beta W^T + bias) in its epilogue (before the GEGLU gate; with the temporal
    sinusoid table as a per-frame bias row).  33000 rows: >= 3 tiles per CTA (the epilogue groups take alternate tiles)
    and a ragged last tile; 1000 / 2048 rows: both groups share every tile."""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(21)
    K0 = 320
    a0 = torch.randn(M, K0, generator=g).half().to(DEV)
    w0 = (torch.randn(C, K0, generator=g) / K0 ** 0.5).half().to(DEV)
    b0 = (0.7 + 0.2 * torch.randn(C, generator=g)).to(DEV)  # row mean comparable to the spread
    res = (torch.randn(M, C, generator=g) * 1.5).half().to(DEV)
    x = torch.empty(M, C, dtype=torch.float16, device=DEV)
    nparts = 3 * (C // 160)
    parts = torch.full((nparts, M, 2), float("nan"), dtype=torch.float32, device=DEV)
    L.gemm([L.Seg(a0, K0, K0, 1)], 1, 1, M, w0, C, x, C, bias=b0, residual=res, ldr=C, tile_n=160, cta_pair=pair,
           row_partials_out=parts)
    xf = x.float()
    want = a0.float() @ w0.float().t() + b0 + res.float()
    assert rel_l2(x, want) < 1e-3
    assert not torch.isnan(parts).any()
    # (the sums are taken before the rounding to fp16: they match the stored values to the rounding noise of C halves)
    assert torch.allclose(parts[..., 0].sum(0), xf.sum(1), atol=5e-2, rtol=1e-4)
    assert torch.allclose(parts[..., 1].sum(0), (xf * xf).sum(1), rtol=5e-4)
    with pytest.raises(RuntimeError):  # the part count is part of the contract
        L.gemm([L.Seg(a0, K0, K0, 1)], 1, 1, M, w0, C, x, C, bias=b0, tile_n=128, row_partials_out=parts)

    w = (torch.randn(N, C, generator=g) / C ** 0.5).to(DEV)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.3 * torch.randn(C, generator=g)).to(DEV)
    ln = F.layer_norm(xf, (C,), gamma, beta, 1e-5)
    flags, tile, n_out, bias_div, bias_ld = 0, 0, N, 0, 0
    if mode == "pe":
        Fr = 16
        hw = M // (2 * Fr)
        pe = torch.randn(Fr, C, generator=g).to(DEV)
        wg, cs, b2 = L.fold_layernorm(w, bias, gamma, beta, pe=pe)
        b2 = b2.repeat(2, 1).contiguous()  # one bias row per (batch element, frame)
        bias_div, bias_ld = hw, N
        ref = (ln.reshape(2, Fr, hw, C) + pe[None, :, None, :]).reshape(M, C) @ w.t() + bias
    else:
        wg, cs, b2 = L.fold_layernorm(w, bias, gamma, beta)
        ref = ln @ w.t() + bias
    if mode == "geglu":
        wg, b2 = L.pack_geglu(wg, b2, 256)
        cs = L.pack_geglu(cs[:, None], None, 256)[0][:, 0].contiguous()
        ref = ref[:, : N // 2] * F.gelu(ref[:, N // 2:])
        n_out, flags, tile = N // 2, L.EPI_GEGLU, 256
    out = torch.empty(M, n_out, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(x, C, C, 1)], 1, 1, M, wg.contiguous(), N, out, n_out, bias=b2, bias_div=bias_div, bias_ld=bias_ld,
           flags=flags, tile_n=tile, cta_pair=pair, col_sum=cs, row_partials_in=parts)
    assert rel_l2(out, ref) < 2e-3
    out2 = torch.empty_like(out)
    L.gemm([L.Seg(x, C, C, 1)], 1, 1, M, wg.contiguous(), N, out2, n_out, bias=b2, bias_div=bias_div, bias_ld=bias_ld,
           flags=flags, tile_n=tile, cta_pair=pair, col_sum=cs, row_partials_in=parts)
    assert torch.equal(out, out2)


def test_layernorm_partials_from_two_launches():
    """the rows of one activation may come from two producer launches (the classifier-free-guidance halves of the audio
    cross-attention, engine.py): part-major layout with an explicit stride; the consumer may also read a row range"""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(23)
    M, C, N = 4096, 640, 640
    half = M // 2
    a0 = torch.randn(M, C, generator=g).half().to(DEV)
    w0 = (torch.randn(C, C, generator=g) / C ** 0.5).half().to(DEV)
    x = torch.empty(M, C, dtype=torch.float16, device=DEV)
    parts = torch.full((3 * (C // 160), M, 2), float("nan"), dtype=torch.float32, device=DEV)
    for e in range(2):
        L.gemm([L.Seg(a0[e * half:], C, C, 1)], 1, 1, half, w0, C, x[e * half:], C, tile_n=160,
               row_partials_out=parts[:, e * half:(e + 1) * half])
    xf = x.float()
    assert torch.allclose(parts[..., 0].sum(0), xf.sum(1), atol=5e-2, rtol=1e-4)
    w = (torch.randn(N, C, generator=g) / C ** 0.5).to(DEV)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.3 * torch.randn(C, generator=g)).to(DEV)
    wg, cs, b2 = L.fold_layernorm(w, None, gamma, beta)
    out = torch.empty(half, N, dtype=torch.float16, device=DEV)
    L.gemm([L.Seg(x[half:], C, C, 1)], 1, 1, half, wg, N, out, N, bias=b2, col_sum=cs,
           row_partials_in=parts[:, half:])
    ref = F.layer_norm(xf[half:], (C,), gamma, beta, 1e-5) @ w.t()
    assert rel_l2(out, ref) < 2e-3


def test_softmax_and_transpose():
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(11)
    s = (torch.randn(512, 1024, generator=g) * 40).to(DEV)
    p = torch.empty(512, 1024, dtype=torch.float16, device=DEV)
    L.softmax_rows(s, 512, 1024, p, scale=0.1)
    assert rel_l2(p, torch.softmax(s * 0.1, -1)) < 2e-3
    x = torch.randn(3, 100, 72, generator=g).half().to(DEV)
    y = torch.empty(3, 72, 100, dtype=torch.float16, device=DEV)
    L.transpose(x, 3, 100, 72, y)
    assert torch.equal(y, x.transpose(1, 2).contiguous())


# ----------------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize(
    "batch,heads,d,sq,skv",
    [(4, 8, 40, 1024, 1024), (4, 8, 80, 256, 256), (4, 8, 160, 64, 64), (3, 8, 160, 16, 16), (4, 8, 40, 1024, 50),
     (4, 8, 80, 256, 50), (2, 8, 160, 64, 50), (2, 8, 160, 16, 50), (2, 8, 40, 100, 77), (1, 8, 40, 4096, 4096),
     # ragged shapes on the tcgen05 path (sq >= 128, skv >= 64): partial query tiles, masked keys in the last tile
     (2, 8, 40, 300, 200), (3, 8, 80, 200, 130), (2, 8, 160, 256, 256), (2, 8, 160, 130, 65), (2, 8, 80, 1024, 1024)],
)
def test_attention(batch, heads, d, sq, skv):
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(12)
    C = heads * d
    q = torch.randn(batch * sq, C, generator=g).half().to(DEV)
    k = torch.randn(batch * skv, C, generator=g).half().to(DEV)
    v = torch.randn(batch * skv, C, generator=g).half().to(DEV)
    out = torch.empty_like(q)
    L.attention(q, k, v, out, C, C, C, C, batch, heads, d, sq, skv)
    qh = (q.float() * d ** -0.5).reshape(batch, sq, heads, d).transpose(1, 2)
    kh = k.float().reshape(batch, skv, heads, d).transpose(1, 2)
    vh = v.float().reshape(batch, skv, heads, d).transpose(1, 2)
    ref = torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh
    ref = ref.transpose(1, 2).reshape(batch * sq, C)
    assert rel_l2(out, ref) < 3e-3


@pytest.mark.parametrize("d,S", [(40, 1024), (80, 256), (160, 256)])
def test_attention_packed_qkv(d, S):
    """spatial self-attention as the UNet plan calls it (engine._transformer): q, k, v are column blocks of one
    [rows, 3C] matrix (leading dimension 3C), the output overwrites a [rows, C] buffer; score magnitudes like the
    network's (|s| up to ~60 before the softmax scale) exercise the lazy rescale of the tcgen05 kernel."""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(15)
    batch, heads = 3, 8
    C = heads * d
    rows = batch * S
    qkv = (torch.randn(rows, 3 * C, generator=g) * 1.7).half().to(DEV)
    out = torch.empty(rows, C, dtype=torch.float16, device=DEV)
    L.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, 3 * C, 3 * C, 3 * C, C, batch, heads, d, S, S)
    x = qkv.float().reshape(batch, S, 3, heads, d).permute(2, 0, 3, 1, 4)  # (3, B, h, S, d)
    ref = torch.softmax(x[0] @ x[1].transpose(-1, -2) * d ** -0.5, -1) @ x[2]
    ref = ref.permute(0, 2, 1, 3).reshape(rows, C)
    assert rel_l2(out, ref) < 3e-3


@pytest.mark.parametrize("d,HW,Fr", [(40, 64, 16), (80, 16, 16), (160, 4, 16), (40, 1024, 16), (40, 36, 8), (32, 10, 4),
                                     (80, 20, 32), (40, 8, 24)])
def test_attention_temporal(d, HW, Fr):
    """VersatileAttention (motion_module.py:262-313): sequences run over the 16 frames at each pixel, addressed by
    stride inside the (b f) x HW token matrix; q, k, v come packed as one [rows, 3C] matrix.  With LS_ATTN_ONE=1 /
    LS_ATTN_TC_ALL=1 (see test_attention_every_shape_on_the_other_kernels) power-of-two frame counts <= 32 take the packed
    tcgen05 kernels (128 / F pixels per tile, incl. tiles that run past the image: HW = 4, 10, 20, 36); 24 frames always
    stay on the warp-level kernel."""
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(13)
    B, heads = 2, 8
    C = heads * d
    rows = B * Fr * HW
    qkv = torch.randn(rows, 3 * C, generator=g).half().to(DEV)
    out = torch.empty(rows, C, dtype=torch.float16, device=DEV)
    addr = (HW, Fr * HW, 1, HW)
    L.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, 3 * C, 3 * C, 3 * C, C, B * HW, heads, d, Fr, Fr,
                q_addr=addr, kv_addr=addr)
    x = qkv.float().reshape(B, Fr, HW, 3, heads, d).permute(3, 0, 2, 4, 1, 5)  # (3, B, HW, h, F, d)
    ref = torch.softmax(x[0] @ x[1].transpose(-1, -2) * d ** -0.5, -1) @ x[2]  # (B, HW, h, F, d)
    ref = ref.permute(0, 3, 1, 2, 4).reshape(rows, C)
    assert rel_l2(out, ref) < 3e-3


# ------------------------------------------------------------------------------------- denoising-loop pointwise
def test_concat13_cfg_ddim_paste():
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(14)
    Fr, H, W = 16, 32, 32
    HW = H * W
    lat = torch.randn(1, 4, Fr, H, W, generator=g).to(DEV)
    mask = (torch.rand(1, 1, Fr, H, W, generator=g) > 0.5).float().to(DEV)
    masked = torch.randn(1, 4, Fr, H, W, generator=g).to(DEV)
    ref = torch.randn(1, 4, Fr, H, W, generator=g).to(DEV)
    out = torch.empty(2 * Fr * HW, 64, dtype=torch.float16, device=DEV)
    L.concat13(lat, mask, masked, ref, 2, Fr, HW, out)
    x = torch.cat([lat, mask, masked, ref], dim=1)
    x = torch.cat([x] * 2)  # (2, 13, F, H, W)
    want = x.permute(0, 2, 3, 4, 1).reshape(2 * Fr * HW, 13)
    assert torch.equal(out[:, :13], want.half())
    assert torch.count_nonzero(out[:, 13:]) == 0

    # CFG combine + DDIM update
    eps_cl = torch.randn(2 * Fr * HW, 32, generator=g).to(DEV)
    lat0 = lat.clone()
    eps_out = torch.empty(1, 4, Fr, H, W, device=DEV)
    gs, a_t, a_p = 1.5, 0.0081550, 0.0120
    L.cfg_ddim_step(eps_cl, 32, 2, Fr, HW, gs, a_t, a_p, lat, eps_out)
    e = eps_cl[:, :4].reshape(2, Fr, H, W, 4).permute(0, 4, 1, 2, 3)
    eps = e[0:1] + gs * (e[1:2] - e[0:1])
    x0 = (lat0 - math.sqrt(1 - a_t) * eps) / math.sqrt(a_t)
    want = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
    assert torch.allclose(eps_out, eps, rtol=1e-6, atol=1e-6)
    assert rel_l2(lat, want) < 1e-6

    # paste-back
    n, HWp = 4, 64 * 64
    dec = torch.randn(n * HWp, 32, generator=g).to(DEV)
    refp = torch.randn(n, 3, HWp, generator=g).to(DEV)
    m = (torch.rand(n, 1, HWp, generator=g) > 0.4).float().to(DEV)
    o = torch.empty(n, 3, HWp, device=DEV)
    L.paste_back(dec, 32, refp, m, n, HWp, o)
    d = dec[:, :3].reshape(n, HWp, 3).permute(0, 2, 1)
    assert torch.allclose(o, d * (1 - m) + refp * m, rtol=1e-6, atol=1e-6)


def test_layout_and_time_embedding():
    L = _ops()
    g = torch.Generator(device="cpu").manual_seed(15)
    B, C, Fr, HW = 2, 13, 4, 64
    x = torch.randn(B, C, Fr, HW, generator=g).to(DEV)
    out = torch.empty(B * Fr * HW, 64, dtype=torch.float16, device=DEV)
    L.ncfhw_to_cl(x, B, C, Fr, HW, 64, 0.5, out)
    want = (x * 0.5).permute(0, 2, 3, 1).reshape(B * Fr * HW, C).half()
    assert torch.equal(out[:, :C], want) and torch.count_nonzero(out[:, C:]) == 0
    y = torch.randn(B * Fr * HW, 32, generator=g).to(DEV)
    back = torch.empty(B, 4, Fr, HW, device=DEV)
    L.cl_to_ncfhw(y, 32, B, 4, Fr, HW, back)
    assert torch.equal(back, y[:, :4].reshape(B, Fr, HW, 4).permute(0, 3, 1, 2))

    t = torch.tensor([951.0, 1.0], device=DEV)
    emb = torch.empty(2, 320, device=DEV)
    L.timestep_embedding(t, 2, 320, emb)
    k = torch.arange(160, device=DEV, dtype=torch.float32)
    freq = torch.exp(-math.log(10000.0) * k / 160)
    e = t[:, None] * freq[None]
    want = torch.cat([torch.cos(e), torch.sin(e)], -1)
    assert torch.allclose(emb, want, atol=2e-4)

    w = (torch.randn(1280, 320, generator=g) / math.sqrt(320)).half().to(DEV)
    b = torch.randn(1280, generator=g).to(DEV)
    add = torch.randn(1280, generator=g).to(DEV)
    yv = torch.empty(2, 1280, device=DEV)
    L.small_linear(emb, 2, 320, w, b, add, 1280, True, True, yv)
    want = F.silu(F.silu(emb) @ w.float().t() + b) + add
    assert rel_l2(yv, want) < 1e-5


@pytest.mark.parametrize("switch", ["LS_ATTN_TC_ALL=1", "LS_ATTN_ONE=1"])
def test_attention_every_shape_on_the_other_kernels(switch):
    """The one-key-tile problems (audio cross-attention, 8x8 / 4x4 levels, temporal attention) run on the warp-level
    kernels of attention.cu by default (they are faster there).  LS_ATTN_ONE=1 routes them to the persistent one-tile
    tcgen05 kernel, LS_ATTN_TC_ALL=1 to the tcgen05 flash kernel (packed temporal mode: block-diagonal mask over two key
    tiles): both stay correct.
    The switches are read once per process, so the attention tests re-run in a child process."""
    import subprocess
    import sys

    name, val = switch.split("=")
    env = dict(os.environ, **{name: val})
    res = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-k",
                          "attention and not every_shape", "--no-header", "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_gemm_weight_resident_mode():
    """LS_GEMM_BRES=1 (off by default - it measured no faster, profiles/r2l_gemm_epilogue_experiments.txt (9)): a CTA of a
    short-K launch keeps its weight tile in shared memory across its M tiles and streams only activations.  Same results as
    an fp32 matmul on every shape that takes the mode (bias, residual, GEGLU, ragged M).  The switch is read once per
    process, hence the child."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "bres_check.py")], env=dict(os.environ, LS_GEMM_BRES="1"),
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("M=")]
    assert len(lines) >= 6 and all(l.rstrip().endswith("OK") for l in lines), res.stdout
