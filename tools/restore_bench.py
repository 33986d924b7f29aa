"""16 x 1080p paste-back (bench.py's restore leg alone): device-resident ms per call; run under ncu for the launch list"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from latentsync_b200.restore import FaceRestorer
from oracle import restore_ref as RR
dev = "cuda"
F = 16
cases = [RR.synthetic_case(500 + i, 1080, 1920, (0.45, 0.6), (500.0, 700.0)) for i in range(F)]
frames = torch.from_numpy(np.stack([c[0] for c in cases])).to(dev)
faces = torch.from_numpy(np.stack([c[1] for c in cases])).to(dev)
mats = [c[2] for c in cases]
out = torch.empty_like(frames)
r = FaceRestorer(dev)
for _ in range(3):
    r.restore_imgs(frames, faces, mats, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("REPS", "10"))
a.record()
for _ in range(n):
    r.restore_imgs(frames, faces, mats, out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"restore 16 x 1080p: {ms:.3f} ms per call, {F / ms * 1e3:.0f} frames/s")
if os.environ.get("CHECK", "0") == "1":
    o = out.cpu().numpy()
    for i in (0, 7, 15):
        assert np.array_equal(o[i], RR.restore_img_cv2(*cases[i])), i
    print("byte-exact vs OpenCV on frames 0, 7, 15")
