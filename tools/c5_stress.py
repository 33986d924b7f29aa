"""BASELINE.json configs[4]: 512 x 512 stress shape (UNet input 13 x 16 x 64 x 64, temporal layers on), a batch of
segments, forward only - conv / attention throughput at that size (stage2 weights, synthetic)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from latentsync_b200 import synthetic as syn
from latentsync_b200.engine import UNetEngine
from latentsync_b200.spec import STAGE2_UNET_CONFIG

nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.cuda.set_device(0)
cfg = STAGE2_UNET_CONFIG
eng = UNetEngine({k: v.cuda() for k, v in syn.unet_state_dict(cfg, 0).items()}, cfg, "cuda")
plan = eng.plan(2 * nseg, 16, 64, 64, 50)
plan.x_in.tensor().normal_()
plan.audio_in.tensor().normal_()
plan.t_in.tensor().fill_(501.0)
plan.capture()
plan.run_hoisted()  # time path + audio K/V: outside the graph
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.replay()
a.record()
for _ in range(3):
    plan.replay()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
kinds = {k: plan.time_kind_in_graph(k, reps=2) for k in ("gemm", "attention", "groupnorm", "layernorm")}
out = {"config": f"configs[4]: {nseg} segments x CFG 2, 13x16x64x64 input, forward only", "forward_ms": ms,
       "flops": plan.flops(), "tflops": plan.flops() / (ms * 1e-3) / 1e12,
       "gemm_tflops": plan.flops("gemm") / (kinds["gemm"] * 1e-3) / 1e12,
       "attention_tflops": plan.flops("attention") / (kinds["attention"] * 1e-3) / 1e12,
       "kind_ms_in_graph": kinds, "launches": plan.launches, "pool_GB": plan.pool.total_bytes() / 1e9}
print(json.dumps(out))
