"""Is the GEMM main loop bound per SM or by the chip's L2 -> SM fabric?  Same tile (128 x BN x 64 k-blocks, L2-resident
operands), launched on 37 / 74 / 148 SMs (one tile per CTA): clocks per k-block and bytes per clock per SM / per chip."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 10
def t(M, N, K, bn):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    outs = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(2)]
    def launch(i):
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, outs[i % 2], N, tile_n=bn, cta_pair=1)
    launch(0); launch(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS): launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * REPS)
mhz = float(os.environ.get("SM_MHZ", "1935"))
for bn in (64, 160, 256):
    for (mt, ntl) in ((1, 37), (2, 37), (4, 37), (37, 1), (37, 4), (74, 2)):
        M, N = 128 * mt, bn * ntl
        ctas = mt * ntl
        t1, t2 = t(M, N, 3200, bn), t(M, N, 6400, bn)
        clk = (t2 - t1) * mhz / 50
        byt = (128 + bn) * 128
        print(f"BN={bn:3d} {mt:3d} M-tiles x {ntl:3d} N-tiles = {ctas:3d} CTAs: {clk:6.0f} clk / k-block (tensor {2*bn}), "
              f"{byt/clk:5.1f} B/clk/SM, {byt*ctas/clk:6.0f} B/clk chip", flush=True)
