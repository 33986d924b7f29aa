"""Host side of the GPU inverse-affine paste-back.

Mirrors `AlignRestore.restore_img(input_img, face, affine_matrix)` (latentsync/utils/affine_transform.py:85-115) for a
batch of frames: `FaceRestorer.restore_imgs(frames, faces, affine_matrices)` returns the same uint8 frames the
reference produces one by one with OpenCV on the CPU.  The arithmetic lives in `csrc/restore.cu`
(`ls_restore_faces`); this file builds what the kernels need from the 2 x 3 matrices: OpenCV's two double-precision
matrix inversions, the frame rectangle the face can touch (ROI), and OpenCV's interpolation tables (third-party
`opencv-python`, 4.13.0 in this image - algorithm restated, no cv2 call on this path).  There is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from . import _lib as L

FACE_SIZE = (210, 280)  # AlignRestore.face_size (w, h) = (int(75 * 2.8), int(100 * 2.8)), affine_transform.py:40-43
GMAX = 128  # largest w_edge with a precomputed Gaussian kernel (w_edge = sqrt(face area in the frame) // 20)


def invert_affine(m: np.ndarray) -> np.ndarray:
    """cv2.invertAffineTransform / the inversion inside cv::warpAffine, in the same double-precision operation order
    (affine_transform.py:89 and the warpAffine calls at :96,:98, which are made WITHOUT WARP_INVERSE_MAP)."""
    m = np.asarray(m, np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def lanczos4_table() -> np.ndarray:
    """OpenCV's fixed-point Lanczos-4 table for 8-bit remap (imgwarp.cpp initInterTab2D / interpolateLanczos4):
    [fy][fx][ky][kx] int16, every 8 x 8 entry sums to 32768 (the rounding residue goes to the largest / smallest of
    the four central taps)."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    t1 = np.zeros((32, 8), np.float32)
    for t in range(32):
        x = np.float32(t) * np.float32(1.0 / 32)
        if x < np.float32(1.1920929e-07):
            t1[t, 3] = 1
            continue
        y0 = -(float(x) + 3) * math.pi * 0.25
        s0, c0 = math.sin(y0), math.cos(y0)
        c = np.zeros(8, np.float32)
        acc = np.float32(0)
        for i in range(8):
            y = -(float(x) + 3 - i) * math.pi * 0.25
            c[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
            acc = np.float32(acc + c[i])
        t1[t] = c * (np.float32(1.0) / acc)
    out = np.zeros((32, 32, 8, 8), np.int16)
    for i in range(32):
        for j in range(32):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(32768)), -32768, 32767).astype(np.int64)
            diff = int(it.sum()) - 32768
            if diff != 0:
                sub = it[4:6, 4:6]
                k = np.unravel_index(np.argmax(sub) if diff < 0 else np.argmin(sub), sub.shape)
                it[4 + k[0], 4 + k[1]] -= diff
            out[i, j] = it
    return out


def gaussian_kernel(n: int) -> np.ndarray:
    """cv2.getGaussianKernel(n, 0, CV_32F) for odd n > 9 and the fixed small kernels below that
    (smooth.dispatch.cpp getGaussianKernelBitExact): sigma = 0.15 n + 0.35, double arithmetic, float result."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
             9: [4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256, 51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256]}
    if n in small:
        return np.asarray(small[n], np.float32)
    sigma = float(n) * 0.15 + 0.35
    scale2x = -0.125 / (sigma * sigma)
    n2 = (n - 1) // 2
    vals = [math.exp(float(x * x) * scale2x) for x in range(1 - n, 1 - n + 2 * n2, 2)]
    s = 0.0
    for v in vals:
        s += v
    mul = 1.0 / (s * 2.0 + 1.0)
    k = np.zeros(n, np.float64)
    for i, v in enumerate(vals):
        k[i] = k[n - 1 - i] = v * mul
    k[n2] = mul
    return k.astype(np.float32)


def gaussian_table(gmax: int = GMAX) -> np.ndarray:
    tab = np.zeros((gmax + 1, 2 * gmax + 1), np.float32)
    tab[0, 0] = 1.0
    for w in range(1, gmax + 1):
        tab[w, : 2 * w + 1] = gaussian_kernel(2 * w + 1)
    return tab


def face_roi(inverse_affine: np.ndarray, wf: int, hf: int, W: int, H: int, margin: int = 3):
    """Frame rectangle (x0, y0, x1, y1) holding every pixel whose source coordinate can fall inside the face
    (+ bilinear / Lanczos reach): the face rectangle [-1, wf] x [-1, hf] mapped to the frame, grown by `margin`."""
    cx = np.array([-1.0, wf, wf, -1.0])
    cy = np.array([-1.0, -1.0, hf, hf])
    x = inverse_affine[0, 0] * cx + inverse_affine[0, 1] * cy + inverse_affine[0, 2]
    y = inverse_affine[1, 0] * cx + inverse_affine[1, 1] * cy + inverse_affine[1, 2]
    x0 = max(0, int(math.floor(x.min())) - margin)
    y0 = max(0, int(math.floor(y.min())) - margin)
    x1 = min(W, int(math.ceil(x.max())) + margin + 1)
    y1 = min(H, int(math.ceil(y.max())) + margin + 1)
    if x1 <= x0 or y1 <= y0:
        return 0, 0, 0, 0
    return x0, y0, x1, y1


class FaceRestorer:
    """Batched GPU `AlignRestore.restore_img` (upscale_factor 1, the only value the reference constructs:
    affine_transform.py:38-39)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("latentsync_b200.FaceRestorer needs a CUDA device (no CPU fallback)")
        self.lanczos = torch.from_numpy(lanczos4_table()).to(self.device)
        self.gauss = torch.from_numpy(gaussian_table(GMAX)).to(self.device)

    def plan(self, affine_matrices: Sequence[np.ndarray], wf: int, hf: int, W: int, H: int):
        """host tables for a batch: dst->src matrices (F, 6) float64, ROIs (F, 4) int32, max ROI width / height"""
        mats = np.zeros((len(affine_matrices), 6), np.float64)
        rois = np.zeros((len(affine_matrices), 4), np.int32)
        for i, a in enumerate(affine_matrices):
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            inv = invert_affine(np.asarray(a, np.float64).reshape(2, 3))  # affine_transform.py:89 (upscale_factor == 1)
            mats[i] = invert_affine(inv).reshape(6)                        # cv::warpAffine's own inversion
            rois[i] = face_roi(inv, max(wf, FACE_SIZE[0]), max(hf, FACE_SIZE[1]), W, H)
        rw = int(max(1, (rois[:, 2] - rois[:, 0]).max()))
        rh = int(max(1, (rois[:, 3] - rois[:, 1]).max()))
        return mats, rois, rw, rh

    @torch.no_grad()
    def restore_imgs(self, frames, faces, affine_matrices, out=None) -> torch.Tensor:
        """frames: uint8 (F, H, W, 3) numpy array or tensor (host or device); faces: uint8 (F, hf, wf, 3) tensor or
        array; affine_matrices: F matrices (2, 3) as returned by AlignRestore.align_warp_face.  Returns the restored
        frames as a uint8 device tensor (F, H, W, 3)."""
        fr = torch.as_tensor(frames) if not isinstance(frames, torch.Tensor) else frames
        fc = torch.as_tensor(faces) if not isinstance(faces, torch.Tensor) else faces
        if fr.dtype != torch.uint8 or fc.dtype != torch.uint8 or fr.dim() != 4 or fc.dim() != 4 or fr.shape[-1] != 3:
            raise ValueError("restore_imgs expects uint8 (F, H, W, 3) frames and (F, hf, wf, 3) faces")
        if fr.shape[0] != fc.shape[0] or fr.shape[0] != len(affine_matrices):
            raise ValueError("restore_imgs: one face and one affine matrix per frame")
        fr = fr.to(self.device, non_blocking=True).contiguous()
        fc = fc.to(self.device, non_blocking=True).contiguous()
        F, H, W, _ = fr.shape
        hf, wf = fc.shape[1], fc.shape[2]
        mats, rois, rw, rh = self.plan(affine_matrices, wf, hf, W, H)
        mats_d = torch.from_numpy(mats).to(self.device)
        rois_d = torch.from_numpy(rois).to(self.device)
        work = torch.empty(3 * F * rw * rh, dtype=torch.float32, device=self.device)
        scratch = torch.empty(3 * F + 1, dtype=torch.int32, device=self.device)
        if out is None:
            out = torch.empty_like(fr)
        L.restore_faces(fr, out, fc, mats_d, rois_d, self.lanczos, self.gauss, work, scratch, rw, rh, GMAX,
                        mh=FACE_SIZE[1], mw=FACE_SIZE[0])  # the mask is ones(face_size) whatever the face (:97)
        self._last_scratch = scratch
        return out

    def check_status(self) -> None:
        """raises if the last call met a face too large for the Gaussian table (synchronises)"""
        if int(self._last_scratch[-1].item()) != 0:
            raise RuntimeError(f"FaceRestorer: w_edge > {GMAX} (face wider than ~{20 * GMAX} px in the frame)")
