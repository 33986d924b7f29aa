"""UNet step time (CUDA-graph replay of one CFG-batched forward, stage2 config) - for A/B runs on one box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import synthetic as syn
from latentsync_b200.engine import UNetEngine
from latentsync_b200.spec import STAGE2_UNET_CONFIG as CFG
torch.cuda.set_device(0)
eng = UNetEngine({k: v.cuda() for k, v in syn.unet_state_dict(CFG, 0).items()}, CFG, "cuda")
B = int(os.environ.get("LS_B", "2"))
plan = eng.plan(B, 16, 32, 32, 50)
plan.x_in.tensor().normal_(); plan.audio_in.tensor().normal_(); plan.t_in.tensor().fill_(951.0)
plan.capture()
for _ in range(5): plan.replay()
torch.cuda.synchronize()
ts = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): plan.replay()
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 20)
print(f"UNET_STEP_MS B={B} " + " ".join(f"{t:.3f}" for t in ts), {k: os.environ[k] for k in os.environ if k.startswith("LS_")})
