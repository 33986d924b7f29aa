"""does a second instantiation of the UNet step graph, replayed alternately, hide the launch set-up of a 558-node graph?
(a graph exec cannot overlap with itself: its next launch is prepared only after the previous one has finished)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import synthetic as syn
from latentsync_b200.engine import UNetEngine
from latentsync_b200.spec import STAGE2_UNET_CONFIG as cfg

torch.cuda.set_device(0)
eng = UNetEngine({k: v.cuda() for k, v in syn.unet_state_dict(cfg, 0).items()}, cfg, "cuda")
plan = eng.plan(2, 16, 32, 32, 50, uncond_zero=True, same_sample=True)
plan.x_in.tensor().normal_(); plan.audio_in.tensor().normal_(); plan.t_in.tensor().fill_(951.0)
plan.capture()
g0 = plan.graph
g1 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g1):
    plan.run(hoisted=False)
small = torch.zeros(1024, device="cuda")


def timed(fn, n=40):
    fn(0); fn(1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(f"same exec back to back            : {timed(lambda i: g0.replay()):.3f} ms")
print(f"two execs alternating             : {timed(lambda i: (g0 if i & 1 else g1).replay()):.3f} ms")
print(f"same exec + 4 tiny kernels between: {timed(lambda i: (g0.replay(), small.add_(1), small.add_(1), small.add_(1), small.add_(1))):.3f} ms")
print(f"two execs + 4 tiny kernels between: {timed(lambda i: ((g0 if i & 1 else g1).replay(), small.add_(1), small.add_(1), small.add_(1), small.add_(1))):.3f} ms")
