"""Main-loop ablation of the tcgen05 GEMM: clocks per 64-wide k-block with the A loads, the B loads or the MMA issue
removed (LS_DBG_* flags; results are garbage, only time is read).  L2-resident operands (M = 2048, N = 18 tiles), two K
values; the per-k-block figure is the time difference divided by the extra k-blocks of the two waves."""
import sys, os, math
os.environ.setdefault("LS_SO_NAME", "_C_ablate.so")  # make -C latentsync_b200/csrc gemm_ablate
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 10
NO_A, NO_B, NO_MMA = 256, 512, 1024
def t(M, N, K, bn, flags, pair=1):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    outs = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(2)]
    def launch(i):
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, outs[i % 2], N, flags=flags, tile_n=bn, cta_pair=pair)
    launch(0); launch(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS): launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * REPS)
clk_mhz = float(os.environ.get("SM_MHZ", "1935"))
M = 2048
print(f"clocks per k-block (tile 128 x BN x 64; tensor work = 2 BN clk), assuming {clk_mhz:.0f} MHz")
for bn in ((160,) if os.environ.get('QUICK') else (64, 128, 160, 256)):
    N = bn * 18
    row = []
    for name, fl in (("full", 0), ("no A", NO_A), ("no B", NO_B), ("no A,B", NO_A | NO_B), ("no MMA", NO_MMA), ("TMA A only", NO_B | NO_MMA), ("TMA B only", NO_A | NO_MMA), ("nothing", NO_A | NO_B | NO_MMA)):
        t1, t2 = t(M, N, 3200, bn, fl), t(M, N, 6400, bn, fl)
        kb = 2 * 50  # two waves x 50 extra k-blocks
        row.append(f"{name} {(t2 - t1) * clk_mhz / kb:5.0f}")
    print(f"BN={bn:3d}: " + " | ".join(row), flush=True)
if os.environ.get("PAIR", "1") == "1":
    for bn in (128, 160, 192, 256):
        N = bn * 18
        t1, t2 = t(M, N, 3200, bn, 0, pair=2), t(M, N, 6400, bn, 0, pair=2)
        print(f"pair BN={bn}: full {(t2 - t1) * clk_mhz / 100:5.0f} clk per k-block per CTA pair (256 x BN x 64; tensor work 2 BN clk per CTA)", flush=True)
