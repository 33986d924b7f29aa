"""CPU tests of the paste-back oracle (oracle/restore_ref.py) and of the host tables of latentsync_b200/restore.py:
the golden vectors come from the reference's own AlignRestore.restore_img (oracle/make_golden_restore.py)."""
import os

import numpy as np
import pytest

from latentsync_b200 import restore as R
from oracle import restore_ref as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "restore_golden.npz")


def golden_cases():
    g = np.load(GOLD)
    for key in sorted(k for k in g.files if k.startswith("case_")):
        seed, H, W, s0, s1, o0, o1 = g[key]
        frame, face, A = O.synthetic_case(int(seed), int(H), int(W), (s0, s1), (o0, o1))
        yield int(seed), frame, face, A, g[f"out_{int(seed)}"]


def test_numpy_restatement_matches_reference_golden():
    n = 0
    for seed, frame, face, A, ref in golden_cases():
        out = O.restore_img_numpy(frame, face, A)
        assert out.dtype == np.uint8 and np.array_equal(out, ref), f"case {seed}"
        assert (ref != frame).any(), "the case must change pixels"
        n += 1
    assert n >= 5


def test_cv2_statement_matches_reference_golden():
    pytest.importorskip("cv2")
    for seed, frame, face, A, ref in golden_cases():
        assert np.array_equal(O.restore_img_cv2(frame, face, A), ref), f"case {seed}"


def test_host_tables_match_opencv():
    cv2 = pytest.importorskip("cv2")
    tab = R.lanczos4_table()
    assert tab.shape == (32, 32, 8, 8) and (tab.astype(np.int64).sum(axis=(2, 3)) == 32768).all()
    assert tab[0, 0, 3, 3] == 32767 or tab[0, 0, 3, 3] == -32768 or tab[0, 0].astype(np.int64).sum() == 32768
    for n in (1, 3, 5, 7, 9, 11, 17, 41, 101, 2 * R.GMAX + 1):
        assert np.array_equal(cv2.getGaussianKernel(n, 0, cv2.CV_32F).reshape(-1), R.gaussian_kernel(n)), n
    g = R.gaussian_table(8)
    assert g.shape == (9, 17) and g[0, 0] == 1 and np.array_equal(g[3, :7], R.gaussian_kernel(7)) and g[3, 7] == 0
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(2, 3))
        assert np.array_equal(cv2.invertAffineTransform(m), R.invert_affine(m))


def test_lanczos_and_mask_warp_equal_opencv_on_random_matrices():
    cv2 = pytest.importorskip("cv2")
    tab = R.lanczos4_table()
    for seed in range(3):
        frame, face, A = O.synthetic_case(100 + seed, 150, 200, (0.7, 2.5), (-60.0, 80.0))
        inv = R.invert_affine(A)
        M = R.invert_affine(inv)
        assert np.array_equal(cv2.warpAffine(face, inv, (200, 150), flags=cv2.INTER_LANCZOS4),
                              O.warp_lanczos4_u8(face, M, 200, 150, tab))
        assert np.array_equal(cv2.warpAffine(np.ones((O.FACE_H, O.FACE_W), np.float32), inv, (200, 150)),
                              O.warp_linear_ones(M, 200, 150, O.FACE_W, O.FACE_H))


def test_roi_contains_every_changed_pixel():
    for seed, frame, face, A, ref in golden_cases():
        H, W = frame.shape[:2]
        x0, y0, x1, y1 = R.face_roi(R.invert_affine(A), O.FACE_W, O.FACE_H, W, H)
        changed = (ref != frame).any(2)
        ys, xs = np.nonzero(changed)
        assert xs.min() >= x0 and xs.max() < x1 and ys.min() >= y0 and ys.max() < y1, seed
    # a face entirely outside the frame: empty ROI
    A = np.array([[1.0, 0.0, 5000.0], [0.0, 1.0, 5000.0]])
    assert R.face_roi(R.invert_affine(A), O.FACE_W, O.FACE_H, 320, 240) == (0, 0, 0, 0)


def test_restorer_refuses_cpu():
    with pytest.raises(RuntimeError):
        R.FaceRestorer("cpu")
