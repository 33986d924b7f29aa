"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the anti-aliased bilinear resize the reference applies to every
decoded face, `torchvision.transforms.functional.resize(face, (h, w), antialias=True)` at
latentsync/pipelines/lipsync_pipeline.py:350, followed by lines :351-353 (`c h w -> h w c`, `/ 2 + 0.5`, clamp, `* 255`,
`.to(uint8)`).  torchvision forwards to aten::_upsample_bilinear2d_aa (align_corners = False); this file restates that
operator's CPU kernel for float32 tensors, operation by operation, so that the CUDA kernel (ls_resize_aa_u8) can be checked
BYTE for byte:

  * separable: horizontal pass over every input row into a float32 temporary, then the vertical pass (a pass whose size does
    not change is skipped);
  * per output index i of a pass with scale = float32(in) / float32(out):
      center = float32(scale * (i + 0.5)) [product in double], support = scale if scale >= 1 else 1,
      invscale = float32(1 / scale) if scale >= 1 else 1,
      lo = max(int(double(center - support) + 0.5), 0), hi = min(int(double(center + support) + 0.5), in),
      w_j = tri(float32((double(float32(j + lo) - center) + 0.5) * invscale)), tri(x) = 1 - |x| for |x| < 1 else 0,
      then w_j / sum_j w_j (float32 sum in order, float32 division);
  * accumulation order of the x86 builds of PyTorch that have FMA (AVX2 / AVX512 dispatch; checked against
    torch 2.11.0 in tests/test_cpu.py): acc = s_0 w_0, then for the next 4 floor((n - 1) / 4) taps acc = acc + (s_j w_j)
    with separately rounded products (gcc vectorises the loop four taps at a time with an in-order reduction), and for
    the remaining taps acc = fma(s_j, w_j, acc).

Only tests/ may import this module.  fma() is emulated in float64 (the product of two float32 is exact there; the sum is
rounded to 53 bits before the final rounding to 24, which differs from a true fma with probability ~2^-29 per operation).
"""
import numpy as np

f32, f64 = np.float32, np.float64


def aa_weights(in_size: int, out_size: int):
    """per output index: (lo, normalised float32 weights)"""
    scale = f32(in_size) / f32(out_size)
    support = scale if scale >= 1.0 else f32(1.0)
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1.0 else f32(1.0)
    max_interp = int(np.ceil(support)) * 2 + 1
    out = []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        lo = max(int(f64(f32(center - support)) + 0.5), 0)
        hi = min(int(f64(f32(center + support)) + 0.5), in_size)
        size = min(max(hi - lo, 0), max_interp)
        w = np.zeros(size, dtype=f32)
        total = f32(0.0)
        for j in range(size):
            arg = f32((f64(f32(f32(j + lo) - center)) + 0.5) * f64(invscale))
            a = abs(arg)
            w[j] = f32(1.0 - f64(a)) if a < 1.0 else f32(0.0)
            total = f32(total + w[j])
        if total != 0.0:
            w = (w / total).astype(f32)
        out.append((lo, w))
    return out


def _fma(a, b, c):
    return (a.astype(f64) * f64(b) + c.astype(f64)).astype(f32)


def _pass(src: np.ndarray, weights, axis: int) -> np.ndarray:
    """one separable pass along `axis` of a float32 array"""
    src = np.moveaxis(src, axis, -1)
    out = np.empty(src.shape[:-1] + (len(weights),), dtype=f32)
    for o, (lo, w) in enumerate(weights):
        n = len(w)
        acc = (src[..., lo] * w[0]).astype(f32)
        nv = ((n - 1) // 4) * 4
        for j in range(1, n):
            s = src[..., lo + j]
            if j <= nv:
                acc = (acc + (s * w[j]).astype(f32)).astype(f32)
            else:
                acc = _fma(s, w[j], acc)
        out[..., o] = acc
    return np.moveaxis(out, -1, axis)


def resize_aa(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """x: float32 (..., H, W) -> (..., oh, ow)"""
    x = np.ascontiguousarray(x, dtype=f32)
    H, W = x.shape[-2:]
    if ow != W:
        x = _pass(x, aa_weights(W, ow), -1)
    if oh != H:
        x = _pass(x, aa_weights(H, oh), -2)
    return x


def restore_faces_u8(faces: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """faces float32 (f, 3, H, W) in about [-1, 1] -> uint8 (f, oh, ow, 3): lipsync_pipeline.py:350-353"""
    r = resize_aa(faces, oh, ow)
    r = np.transpose(r, (0, 2, 3, 1))
    r = (r / f32(2.0)).astype(f32) + f32(0.5)
    r = np.clip(r.astype(f32), f32(0.0), f32(1.0))
    return (r * f32(255.0)).astype(f32).astype(np.uint8)
