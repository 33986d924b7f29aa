"""the level-0 temporal attention launch (2 x 1024 pixels x 16 frames, 8 heads x 40) for ncu / timing"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
B, Fr, HW, heads, d = 2, 16, int(os.environ.get("HW", 1024)), 8, int(os.environ.get("D", 40))
C = heads * d
rows = B * Fr * HW
qkv = torch.randn(rows, 3 * C, device=dev).half()
out = torch.empty(rows, C, dtype=torch.float16, device=dev)
addr = (HW, Fr * HW, 1, HW)
def run():
    L.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, 3 * C, 3 * C, 3 * C, C, B * HW, heads, d, Fr, Fr, q_addr=addr, kv_addr=addr)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"temporal HW={HW} d={d}: {e0.elapsed_time(e1) * 100:.1f} us per launch (eager, back to back)")
