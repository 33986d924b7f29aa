"""level-0 norm launches for ncu: LayerNorm (32768 x 320), per-frame GroupNorm (32 x 1024 rows), joint GroupNorm + SiLU
(2 x 16384 rows); input written by a preceding elementwise kernel so that it is L2-resident like in the network"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
rows, C = 32768, 320
x = torch.randn(rows, C, device=dev).half()
y = torch.empty_like(x)
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
stats = torch.zeros(1 << 16, device=dev)
def ln(): L.layernorm(x, rows, C, g, b, 1e-5, y)
def gn_frame(): L.groupnorm_fused(x, C, None, 0, rows, 1024, 32, g, b, 1e-6, False, y, stats)
def gn_joint(): L.groupnorm_fused(x, C, None, 0, rows, 16384, 32, g, b, 1e-5, True, y, stats)
for fn in (ln, gn_frame, gn_joint):
    for _ in range(3):
        x.mul_(1.0)  # producer: leaves x in L2
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(10):
            x.mul_(1.0)
            fn()
    gph.replay(); torch.cuda.synchronize()
    e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
    t_both = e0.elapsed_time(e1) * 100
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(10):
            x.mul_(1.0)
    g2.replay(); torch.cuda.synchronize()
    e0.record(); g2.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{fn.__name__}: {t_both - e0.elapsed_time(e1) * 100:.1f} us per launch in a graph behind its producer (producer alone {e0.elapsed_time(e1) * 100:.1f} us)")
