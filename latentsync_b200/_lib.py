"""ctypes binding of the C-ABI in include/latentsync_b200.h (built in-tree as latentsync_b200/_C.so).

There is deliberately no CPU or PyTorch fallback: every op raises if the CUDA extension is missing or a call
fails.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, os.environ.get("LS_SO_NAME", "_C.so"))  # LS_SO_NAME: instrumented debug builds
_CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 2  # include/latentsync_b200.h LS_ABI_VERSION: bumped whenever an argument struct changes
EPI_GEGLU = 1
EPI_OUT_F32 = 2
EPI_SILU = 4
MAX_SEG = 3


class LsGemmArgs(C.Structure):
    _fields_ = [
        ("nseg", C.c_int32),
        ("a_ptr", C.c_void_p * MAX_SEG),
        ("a_ch", C.c_int32 * MAX_SEG),
        ("a_ld", C.c_int32 * MAX_SEG),
        ("a_taps", C.c_int32 * MAX_SEG),
        ("nimg", C.c_int32),
        ("H", C.c_int32),
        ("W", C.c_int32),
        ("b_ptr", C.c_void_p),
        ("N", C.c_int32),
        ("Ktot", C.c_int32),
        ("b_batch_stride", C.c_int64),
        ("bias", C.c_void_p),
        ("bias_div", C.c_int32),
        ("bias_ld", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int32),
        ("out", C.c_void_p),
        ("ldo", C.c_int32),
        ("flags", C.c_int32),
        ("tile_n", C.c_int32),
        ("cta_pair", C.c_int32),
        ("col_sum", C.c_void_p),
        ("row_partials_in", C.c_void_p),
        ("n_partials_in", C.c_int32),
        ("partials_in_stride", C.c_int64),
        ("ln_eps", C.c_float),
        ("row_partials_out", C.c_void_p),
        ("n_partials_out", C.c_int32),
        ("partials_out_stride", C.c_int64),
        ("gn_partials_out", C.c_void_p),
        ("gn_unit", C.c_int32),
        ("gn_partials_ld", C.c_int32),
        ("up2", C.c_int32),
        ("stride2", C.c_int32),
        ("stride2_pad", C.c_int32),
    ]


class LsAttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("out", C.c_void_p),
        ("ldq", C.c_int32),
        ("ldk", C.c_int32),
        ("ldv", C.c_int32),
        ("ldo", C.c_int32),
        ("batch", C.c_int32),
        ("heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("sq", C.c_int32),
        ("skv", C.c_int32),
        ("q_inner", C.c_int32),
        ("q_outer_stride", C.c_int64),
        ("q_inner_stride", C.c_int64),
        ("q_seq_stride", C.c_int64),
        ("kv_inner", C.c_int32),
        ("kv_outer_stride", C.c_int64),
        ("kv_inner_stride", C.c_int64),
        ("kv_seq_stride", C.c_int64),
        ("scale", C.c_float),
    ]


class LsRestoreArgs(C.Structure):
    _fields_ = [
        ("frames", C.c_void_p),
        ("out", C.c_void_p),
        ("faces", C.c_void_p),
        ("mats", C.c_void_p),
        ("rois", C.c_void_p),
        ("lanczos_tab", C.c_void_p),
        ("gauss_tab", C.c_void_p),
        ("work", C.c_void_p),
        ("scratch", C.c_void_p),
        ("F", C.c_int32),
        ("H", C.c_int32),
        ("W", C.c_int32),
        ("hf", C.c_int32),
        ("wf", C.c_int32),
        ("mh", C.c_int32),
        ("mw", C.c_int32),
        ("RW", C.c_int32),
        ("RH", C.c_int32),
        ("gmax", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/latentsync_b200.h declares (tests check this)
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "ls_last_error": (C.c_char_p, []),
    "ls_abi_version": (C.c_int, []),
    "ls_launch_count": (C.c_int64, []),
    "ls_reset_launch_count": (None, []),
    "ls_gemm": (C.c_int, [C.POINTER(LsGemmArgs), _vp]),
    "ls_groupnorm_stats": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "ls_groupnorm_apply": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _f32, _i32, _vp, _vp]),
    "ls_groupnorm": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _f32, _i32, _vp, _vp, _vp]),
    "ls_groupnorm_parts": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _f32, _i32,
                                     _vp, _vp]),
    "ls_layernorm": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _i32, _i32, _vp, _vp]),
    "ls_attention": (C.c_int, [C.POINTER(LsAttnArgs), _vp]),
    "ls_softmax_rows": (C.c_int, [_vp, _i64, _i32, _f32, _vp, _vp]),
    "ls_transpose": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "ls_concat13": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "ls_cfg_ddim_step": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp, _vp]),
    "ls_ncfhw_to_cl": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "ls_cl_to_ncfhw": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_upsample2x": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_im2col_s2": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_im2col_s2_pad": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_gaussian_sample": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp]),
    "ls_preprocess_u8": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "ls_resize_aa_u8": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_paste_back": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "ls_small_linear": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "ls_timestep_embedding": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "ls_fill_zero": (C.c_int, [_vp, _i64, _vp]),
    "ls_im2col1d": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ls_whisper_chunks": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    "ls_restore_faces": (C.c_int, [C.POINTER(LsRestoreArgs), _vp]),
}

_lib: Optional[C.CDLL] = None


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into latentsync_b200/_C.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", _CSRC, "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building latentsync_b200/_C.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return _SO


def lib() -> C.CDLL:
    """Load the extension; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"latentsync_b200: CUDA extension {_SO} is missing - run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)"
            )
        loaded = C.CDLL(_SO)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(loaded, name)
            fn.restype = res
            fn.argtypes = args
        if loaded.ls_abi_version() != ABI_VERSION:
            # a stale build would read the argument structs below with another layout: fail loudly instead
            raise RuntimeError(f"latentsync_b200: {_SO} implements ABI v{loaded.ls_abi_version()}, this binding needs "
                               f"v{ABI_VERSION} - rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        _lib = loaded
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().ls_last_error().decode()}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda, "latentsync_b200 ops need CUDA tensors (no CPU fallback)"
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().ls_launch_count())


def reset_launch_count() -> None:
    lib().ls_reset_launch_count()


# ------------------------------------------------------------------------------------------------ op wrappers
class Seg:
    """One K segment of the GEMM A operand: channels-last fp16 tensor, `ch` channels used, 1 or 9 taps."""

    __slots__ = ("t", "ch", "ld", "taps")

    def __init__(self, t: torch.Tensor, ch: int, ld: int, taps: int = 1):
        self.t, self.ch, self.ld, self.taps = t, ch, ld, taps


def gemm(
    segs: Sequence[Seg],
    nimg: int,
    H: int,
    W: int,
    weight: torch.Tensor,
    N: int,
    out: torch.Tensor,
    ldo: int,
    bias: Optional[torch.Tensor] = None,
    bias_div: int = 0,
    bias_ld: int = 0,
    residual: Optional[torch.Tensor] = None,
    ldr: int = 0,
    flags: int = 0,
    tile_n: int = 0,
    b_batch_stride: int = 0,
    cta_pair: int = 0,
    col_sum: Optional[torch.Tensor] = None,
    row_partials_in: Optional[torch.Tensor] = None,
    row_partials_out: Optional[torch.Tensor] = None,
    ln_eps: float = 1e-5,
    gn_partials_out: Optional[torch.Tensor] = None,
    gn_unit: int = 0,
    up2: int = 0,
    stride2: int = 0,
    stride2_pad: int = 1,
) -> None:
    """row_partials_in / row_partials_out: fp32 [parts, rows, 2] (LsGemmArgs: LayerNorm folded into the consuming GEMM);
    gn_partials_out: fp32 [rows / 128, N / gn_unit (or more), 2] (GroupNorm statistics from this GEMM's epilogue)"""
    a = LsGemmArgs()
    a.up2 = up2  # 1 + 2 py + px: sub-pixel phase of upsample -> conv (segments with 4 taps, low-resolution geometry)
    a.stride2, a.stride2_pad = stride2, stride2_pad  # stride-2 3x3 conv read in place (output geometry, input tensor)
    if gn_partials_out is not None:
        t = gn_partials_out
        assert t.dtype == torch.float32 and t.dim() == 3 and t.stride(2) == 1 and t.stride(1) == 2 and t.stride(0) % 2 == 0
        a.gn_partials_out, a.gn_unit, a.gn_partials_ld = _ptr(t), gn_unit, t.stride(0) // 2
    a.col_sum = _ptr(col_sum)
    for name, t in (("in", row_partials_in), ("out", row_partials_out)):
        if t is not None:  # [parts, rows, 2], possibly a row range of a larger [parts, all_rows, 2] array
            assert t.dtype == torch.float32 and t.dim() == 3 and t.stride(2) == 1 and t.stride(1) == 2 and t.stride(0) % 2 == 0
            setattr(a, "row_partials_" + name, _ptr(t))
            setattr(a, f"n_partials_{name}", t.shape[0])
            setattr(a, f"partials_{name}_stride", t.stride(0) // 2)
    a.ln_eps = ln_eps
    a.nseg = len(segs)
    ktot = 0
    for i, s in enumerate(segs):
        assert s.t.dtype == torch.float16
        a.a_ptr[i] = _ptr(s.t)
        a.a_ch[i] = s.ch
        a.a_ld[i] = s.ld
        a.a_taps[i] = s.taps
        ktot += s.ch * s.taps
    a.nimg, a.H, a.W = nimg, H, W
    assert weight.dtype == torch.float16
    a.b_ptr = _ptr(weight)
    a.N, a.Ktot = N, ktot
    a.b_batch_stride = b_batch_stride
    if bias is not None:
        assert bias.dtype == torch.float32
    a.bias = _ptr(bias)
    a.bias_div = bias_div
    a.bias_ld = bias_ld
    a.residual = _ptr(residual)
    a.ldr = ldr
    a.out = _ptr(out)
    a.ldo = ldo
    a.flags = flags
    a.tile_n = tile_n
    a.cta_pair = cta_pair
    _check(lib().ls_gemm(C.byref(a), _stream()), "ls_gemm")


def groupnorm(
    x1: torch.Tensor,
    c1: int,
    x2: Optional[torch.Tensor],
    c2: int,
    rows: int,
    rows_per_inst: int,
    groups: int,
    gamma: torch.Tensor,
    beta: torch.Tensor,
    eps: float,
    silu: bool,
    out: torch.Tensor,
    stats: torch.Tensor,
) -> None:
    L = lib()
    ninst = rows // rows_per_inst
    assert stats.dtype == torch.float32 and stats.numel() >= ninst * groups * 2
    st = _stream()
    _check(L.ls_groupnorm_stats(_ptr(x1), c1, _ptr(x2), c2, rows, rows_per_inst, groups, _ptr(stats), st),
           "ls_groupnorm_stats")
    _check(
        L.ls_groupnorm_apply(_ptr(x1), c1, _ptr(x2), c2, rows, rows_per_inst, groups, _ptr(stats), _ptr(gamma),
                             _ptr(beta), eps, int(silu), _ptr(out), st),
        "ls_groupnorm_apply",
    )


def groupnorm_fused(x1, c1, x2, c2, rows, rows_per_inst, groups, gamma, beta, eps, silu, out, stats) -> None:
    _check(
        lib().ls_groupnorm(_ptr(x1), c1, _ptr(x2), c2, rows, rows_per_inst, groups, _ptr(gamma), _ptr(beta), eps,
                           int(silu), _ptr(stats), _ptr(out), _stream()),
        "ls_groupnorm",
    )


def groupnorm_parts(x1, c1, parts1, x2, c2, parts2, rows, rows_per_inst, groups, unit, gamma, beta, eps, silu, out) -> None:
    """parts: fp32 [rows / 128, >= c / unit, 2] written by the GEMM(s) that produced x (gemm(gn_partials_out=...))"""
    ld1 = parts1.stride(0) // 2
    ld2 = parts2.stride(0) // 2 if parts2 is not None else 0
    _check(
        lib().ls_groupnorm_parts(_ptr(x1), c1, _ptr(parts1), ld1, _ptr(x2), c2, _ptr(parts2), ld2, rows, rows_per_inst,
                                 groups, unit, _ptr(gamma), _ptr(beta), eps, int(silu), _ptr(out), _stream()),
        "ls_groupnorm_parts",
    )


def layernorm(x, rows, Cc, gamma, beta, eps, out, pe=None, rows_per_frame=1, nframes=1) -> None:
    _check(
        lib().ls_layernorm(_ptr(x), rows, Cc, _ptr(gamma), _ptr(beta), eps, _ptr(pe), rows_per_frame, nframes,
                           _ptr(out), _stream()),
        "ls_layernorm",
    )


def attention(q, k, v, out, ldq, ldk, ldv, ldo, batch, heads, head_dim, sq, skv, q_addr=None, kv_addr=None,
              scale=None) -> None:
    """q_addr / kv_addr = (inner, outer_stride, inner_stride, seq_stride); default = contiguous sequences."""
    a = LsAttnArgs()
    a.q, a.k, a.v, a.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    a.ldq, a.ldk, a.ldv, a.ldo = ldq, ldk, ldv, ldo
    a.batch, a.heads, a.head_dim, a.sq, a.skv = batch, heads, head_dim, sq, skv
    qi = q_addr or (1, sq, 0, 1)
    ki = kv_addr or (1, skv, 0, 1)
    a.q_inner, a.q_outer_stride, a.q_inner_stride, a.q_seq_stride = qi
    a.kv_inner, a.kv_outer_stride, a.kv_inner_stride, a.kv_seq_stride = ki
    a.scale = float(head_dim) ** -0.5 if scale is None else scale
    _check(lib().ls_attention(C.byref(a), _stream()), "ls_attention")


def softmax_rows(s, rows, cols, p, scale=1.0) -> None:
    assert s.dtype == torch.float32
    _check(lib().ls_softmax_rows(_ptr(s), rows, cols, scale, _ptr(p), _stream()), "ls_softmax_rows")


def transpose(x, batch, R, Cc, y) -> None:
    _check(lib().ls_transpose(_ptr(x), batch, R, Cc, _ptr(y), _stream()), "ls_transpose")


def concat13(latents, mask, masked, ref, nb, F, HW, out) -> None:
    for t in (latents, mask, masked, ref):
        assert t.dtype == torch.float32 and t.is_contiguous()
    _check(lib().ls_concat13(_ptr(latents), _ptr(mask), _ptr(masked), _ptr(ref), nb, F, HW, _ptr(out), _stream()),
           "ls_concat13")


def cfg_ddim_step(eps_cl, ld_eps, nb, F, HW, guidance, alpha_t, alpha_prev, latents, eps_out=None) -> None:
    _check(
        lib().ls_cfg_ddim_step(_ptr(eps_cl), ld_eps, nb, F, HW, guidance, alpha_t, alpha_prev, _ptr(latents),
                               _ptr(eps_out), _stream()),
        "ls_cfg_ddim_step",
    )


def ncfhw_to_cl(x, B, Cc, F, HW, cpad, scale, out) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous()
    _check(lib().ls_ncfhw_to_cl(_ptr(x), B, Cc, F, HW, cpad, scale, _ptr(out), _stream()), "ls_ncfhw_to_cl")


def cl_to_ncfhw(x, ld, B, Cc, F, HW, out) -> None:
    assert x.dtype == torch.float32 and out.dtype == torch.float32
    _check(lib().ls_cl_to_ncfhw(_ptr(x), ld, B, Cc, F, HW, _ptr(out), _stream()), "ls_cl_to_ncfhw")


def upsample2x(x, nimg, H, W, Cc, y) -> None:
    _check(lib().ls_upsample2x(_ptr(x), nimg, H, W, Cc, _ptr(y), _stream()), "ls_upsample2x")


def im2col_s2(x, nimg, H, W, Cc, y) -> None:
    _check(lib().ls_im2col_s2(_ptr(x), nimg, H, W, Cc, _ptr(y), _stream()), "ls_im2col_s2")


def im2col_s2_pad(x, nimg, H, W, Cc, pad_before, y) -> None:
    _check(lib().ls_im2col_s2_pad(_ptr(x), nimg, H, W, Cc, pad_before, _ptr(y), _stream()), "ls_im2col_s2_pad")


def preprocess_u8(img, mask, pixel, masked) -> None:
    """img uint8 (n,H,W,3) or (n,3,H,W); mask fp32 (1|3,H,W); pixel / masked fp32 (n,3,H,W)"""
    assert img.dtype == torch.uint8 and mask.dtype == torch.float32 and img.is_contiguous() and mask.is_contiguous()
    hwc = int(img.shape[-1] == 3 and img.shape[1] != 3)
    n = img.shape[0]
    H, W = (img.shape[1], img.shape[2]) if hwc else (img.shape[2], img.shape[3])
    _check(lib().ls_preprocess_u8(_ptr(img), n, H, W, hwc, _ptr(mask), mask.shape[0], _ptr(pixel), _ptr(masked),
                                  _stream()), "ls_preprocess_u8")


def resize_aa_u8(x, oh, ow, out) -> None:
    """x fp32 (n,3,H,W) in [-1,1] -> out uint8 (n,oh,ow,3)"""
    assert x.dtype == torch.float32 and out.dtype == torch.uint8 and x.is_contiguous()
    n, _, H, W = x.shape
    _check(lib().ls_resize_aa_u8(_ptr(x), n, H, W, oh, ow, _ptr(out), _stream()), "ls_resize_aa_u8")


def restore_faces(frames, out, faces, mats, rois, lanczos_tab, gauss_tab, work, scratch, RW, RH, gmax, mh=0, mw=0) -> None:
    """frames / out uint8 (F,H,W,3); faces uint8 (F,hf,wf,3); mats float64 (F,6); rois int32 (F,4); see LsRestoreArgs"""
    assert frames.dtype == torch.uint8 and out.dtype == torch.uint8 and faces.dtype == torch.uint8
    assert mats.dtype == torch.float64 and rois.dtype == torch.int32 and lanczos_tab.dtype == torch.int16
    assert gauss_tab.dtype == torch.float32 and work.dtype == torch.float32
    for t in (frames, out, faces, mats, rois, lanczos_tab, gauss_tab, work, scratch):
        assert t.is_contiguous()
    a = LsRestoreArgs()
    a.frames, a.out, a.faces = _ptr(frames), _ptr(out), _ptr(faces)
    a.mats, a.rois = _ptr(mats), _ptr(rois)
    a.lanczos_tab, a.gauss_tab = _ptr(lanczos_tab), _ptr(gauss_tab)
    a.work, a.scratch = _ptr(work), _ptr(scratch)
    a.F, a.H, a.W = frames.shape[0], frames.shape[1], frames.shape[2]
    a.hf, a.wf = faces.shape[1], faces.shape[2]
    a.mh, a.mw = mh, mw
    a.RW, a.RH, a.gmax = RW, RH, gmax
    assert work.numel() >= 3 * a.F * RW * RH and scratch.numel() * scratch.element_size() >= 12 * a.F + 4
    assert gauss_tab.shape == (gmax + 1, 2 * gmax + 1) and lanczos_tab.numel() == 32 * 32 * 64
    _check(lib().ls_restore_faces(C.byref(a), _stream()), "ls_restore_faces")


def gaussian_sample(moments_cl, ld, noise, n, Cc, HW, shift, scale, z) -> None:
    assert moments_cl.dtype == torch.float32 and z.dtype == torch.float32
    _check(lib().ls_gaussian_sample(_ptr(moments_cl), ld, _ptr(noise) if noise is not None else None, n, Cc, HW,
                                    shift, scale, _ptr(z), _stream()), "ls_gaussian_sample")


def paste_back(decoded_cl, ld, ref, mask, n, HW, out) -> None:
    _check(lib().ls_paste_back(_ptr(decoded_cl), ld, _ptr(ref), _ptr(mask), n, HW, _ptr(out), _stream()),
           "ls_paste_back")


def small_linear(x, B, K, W, bias, add, N, silu_in, silu_out, y) -> None:
    _check(
        lib().ls_small_linear(_ptr(x), B, K, _ptr(W), _ptr(bias), _ptr(add), N, int(silu_in), int(silu_out), _ptr(y),
                              _stream()),
        "ls_small_linear",
    )


def timestep_embedding(t, B, dim, out) -> None:
    _check(lib().ls_timestep_embedding(_ptr(t), B, dim, _ptr(out), _stream()), "ls_timestep_embedding")


def fold_layernorm(weight: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                   pe: Optional[torch.Tensor] = None):
    """nn.LayerNorm(gamma, beta) followed by nn.Linear(weight [N, K], bias) as ONE GEMM on the raw activations
    (LsGemmArgs.col_sum): returns (W' = W diag(gamma) as fp16 [N, K], col_sum fp32 [N] of that fp16 operand,
    bias' = beta W^T + bias, fp32 [N]).  With `pe` ([F, K], the temporal sinusoid table added AFTER the norm,
    motion_module.py:232-234) bias' is [F, N]: row f = (beta + pe[f]) W^T + bias."""
    w32 = weight.float()
    wg = (w32 * gamma.float()[None, :]).to(torch.float16)
    col_sum = wg.float().sum(dim=1).contiguous()
    if pe is None:
        b2 = w32 @ beta.float()
    else:
        b2 = (beta.float()[None, :] + pe.float()) @ w32.t()
    if bias is not None:
        b2 = b2 + bias.float()
    return wg.contiguous(), col_sum, b2.contiguous()


def pack_geglu(weight: torch.Tensor, bias: Optional[torch.Tensor], tile_n: int):
    """Re-order the rows of diffusers' GEGLU projection ([value rows | gate rows], attention.py:171) so that every
    `tile_n`-row tile of the GEMM's B operand holds tile_n/2 value rows followed by their tile_n/2 gate rows; the
    LS_EPI_GEGLU epilogue then finds value and gate of one output column in the same TMEM accumulator."""
    two_inner = weight.shape[0]
    inner = two_inner // 2
    half = tile_n // 2
    assert inner % half == 0
    idx = torch.arange(inner, device=weight.device).reshape(inner // half, half)
    order = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    wp = weight[order].contiguous()
    bp = bias[order].contiguous() if bias is not None else None
    return wp, bp
