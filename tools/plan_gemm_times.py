"""In-graph time of every distinct GEMM launch of the UNet plan the pipeline runs (null-audio shortcut + shared CFG prefix):
each launch is captured alone, REPS times back to back, into its own CUDA graph (operands L2-warm: an A/B tool for
epilogue / tiling changes, not a substitute for the in-step numbers of bench.py).

    LS_FOLD_LN=0 python tools/plan_gemm_times.py > gpurun_out/plan_gemm_fold0.txt
"""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import synthetic as syn  # noqa: E402
from latentsync_b200.engine import UNetEngine  # noqa: E402
from latentsync_b200.spec import STAGE2_UNET_CONFIG  # noqa: E402

REPS = 10
cfg = STAGE2_UNET_CONFIG
torch.cuda.set_device(0)
eng = UNetEngine({k: v.cuda() for k, v in syn.unet_state_dict(cfg, 0).items()}, cfg, "cuda")
plan = eng.plan(2, 16, 32, 32, 50, uncond_zero=True, same_sample=True)
plan.x_in.tensor().normal_()
plan.audio_in.tensor().normal_()
plan.t_in.tensor().fill_(951.0)
plan.run()
torch.cuda.synchronize()
groups = OrderedDict()
for i, (k, d) in enumerate(zip(plan.kinds, plan.descs)):
    if k in ("gemm", "layernorm") and i not in plan.hoisted:
        groups.setdefault((k, d), []).append(i)
tot = {}
rows = []
for (k, d), idx in groups.items():
    fn = plan.ops[idx[0]]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1000 / (2 * REPS)
    rows.append((us * len(idx), len(idx), us, k, d))
    tot[k] = tot.get(k, 0.0) + us * len(idx)
for t, n, us, k, d in sorted(rows, reverse=True):
    print(f"{k:9s} x{n:3d} {us:8.2f} us  {t / 1e3:7.3f} ms  {d}")
print("totals (ms):", {k: round(v / 1e3, 3) for k, v in tot.items()})
