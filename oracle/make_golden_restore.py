"""TEST INFRASTRUCTURE - writes tests/golden/restore_golden.npz from the REFERENCE's own AlignRestore.restore_img
(/root/reference/latentsync/utils/affine_transform.py:85-115; imports only numpy + cv2, so it runs here unmodified).
Inputs are not stored: tests rebuild them with oracle.restore_ref.synthetic_case(seed, H, W).
Run in the build container:  python oracle/make_golden_restore.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from latentsync.utils.affine_transform import AlignRestore  # noqa: E402

from oracle.restore_ref import restore_img_numpy, synthetic_case  # noqa: E402

CASES = [  # seed, H, W, scale range, range of the face origin in the frame
    (11, 240, 320, (1.2, 1.6), (10.0, 60.0)),      # whole face inside a small frame (w_edge ~8)
    (12, 180, 320, (0.5, 0.7), (-120.0, -20.0)),   # face larger than the frame: clipped on several sides
    (13, 200, 240, (3.0, 3.6), (20.0, 100.0)),     # small face (w_edge 3: 7-tap blur)
    (14, 120, 160, (6.0, 7.0), (20.0, 60.0)),      # tiny face: w_edge 1 (3-tap blur)
    (15, 120, 160, (20.0, 24.0), (20.0, 60.0)),    # face of ~10 px: w_edge 0 (empty erosion kernel -> 3 x 3, no blur)
]

if __name__ == "__main__":
    restorer = AlignRestore()
    out = {}
    for seed, H, W, sc, sh in CASES:
        frame, face, A = synthetic_case(seed, H, W, sc, sh)
        res = restorer.restore_img(frame, face, A)
        assert res.dtype == np.uint8
        out[f"out_{seed}"] = res
        out[f"case_{seed}"] = np.array([seed, H, W, sc[0], sc[1], sh[0], sh[1]], np.float64)
        mine, w_edge = restore_img_numpy(frame, face, A, True)
        assert np.array_equal(mine, res), f"numpy restatement differs from the reference on case {seed}"
        print("  restatement == reference, w_edge", w_edge)
        print(seed, H, W, "changed pixels", int((res != frame).any(2).sum()))
    path = os.path.join(ROOT, "tests", "golden", "restore_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
