"""GPU parity of ls_restore_faces (through the C-ABI, latentsync_b200.restore.FaceRestorer) with the reference's
AlignRestore.restore_img: golden vectors made by the reference itself, and OpenCV's own calls at full frame sizes.
The bar is byte-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import restore_ref as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "restore_golden.npz")


@pytest.fixture(scope="module")
def restorer():
    from latentsync_b200.restore import FaceRestorer

    return FaceRestorer("cuda")


def test_golden_from_reference(restorer):
    g = np.load(GOLD)
    keys = sorted(k for k in g.files if k.startswith("case_"))
    assert len(keys) >= 5
    for key in keys:
        seed, H, W, s0, s1, o0, o1 = g[key]
        frame, face, A = O.synthetic_case(int(seed), int(H), int(W), (s0, s1), (o0, o1))
        out = restorer.restore_imgs(frame[None], face[None], [A]).cpu().numpy()[0]
        restorer.check_status()
        ref = g[f"out_{int(seed)}"]
        assert np.array_equal(out, ref), f"case {int(seed)}: {(out != ref).sum()} bytes differ, max " \
                                         f"{np.abs(out.astype(int) - ref.astype(int)).max()}"


def test_batch_of_frames_vs_opencv_720p(restorer):
    pytest.importorskip("cv2")
    cases = [O.synthetic_case(200 + i, 720, 1280, (0.5, 1.3), (100.0, 500.0)) for i in range(6)]
    frames = np.stack([c[0] for c in cases])
    faces = np.stack([c[1] for c in cases])
    out = restorer.restore_imgs(frames, faces, [c[2] for c in cases]).cpu().numpy()
    restorer.check_status()
    for i, (frame, face, A) in enumerate(cases):
        ref = O.restore_img_cv2(frame, face, A)
        assert (ref != frame).any()
        assert np.array_equal(out[i], ref), f"frame {i}: {(out[i] != ref).sum()} bytes differ"


def test_1080p_clipped_faces_and_in_place(restorer):
    pytest.importorskip("cv2")
    cases = [O.synthetic_case(300, 1080, 1920, (0.35, 0.45), (-200.0, 100.0)),   # ~550 px face hanging over the corner
             O.synthetic_case(301, 1080, 1920, (0.6, 0.8), (900.0, 1000.0)),     # lower edge
             O.synthetic_case(302, 1080, 1920, (0.9, 1.0), (5000.0, 6000.0))]    # outside the frame: nothing changes
    frames = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    faces = np.stack([c[1] for c in cases])
    res = restorer.restore_imgs(frames, faces, [c[2] for c in cases], out=frames)  # in place
    restorer.check_status()
    assert res.data_ptr() == frames.data_ptr()
    out = res.cpu().numpy()
    for i, (frame, face, A) in enumerate(cases):
        ref = O.restore_img_cv2(frame, face, A)
        assert np.array_equal(out[i], ref), f"frame {i}: {(out[i] != ref).sum()} bytes differ"
    assert np.array_equal(out[2], cases[2][0])


def test_bad_arguments_raise(restorer):
    frame, face, A = O.synthetic_case(1, 64, 64)
    with pytest.raises(ValueError):
        restorer.restore_imgs(frame[None].astype(np.float32), face[None], [A])
    with pytest.raises(ValueError):
        restorer.restore_imgs(frame[None], face[None], [A, A])
