"""what does the epilogue of a short-K linear cost?  M = 32768 / 8192 / 2048, N = K = C with and without bias / residual"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
def t(M, N, K, bias, res, nset=3, reps=9):
    a = [torch.randn(M, K, device=dev).half() for _ in range(nset)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    b = torch.randn(N, device=dev) if bias else None
    r = [torch.randn(M, N, device=dev).half() for _ in range(nset)] if res else None
    o = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(nset)]
    def launch(i):
        L.gemm([L.Seg(a[i % nset], K, K, 1)], 1, 1, M, w, N, o[i % nset], N, bias=b, residual=r[i % nset] if res else None, ldr=N)
    launch(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * reps)
for (M, C) in ((32768, 320), (8192, 640), (2048, 1280)):
    for nset in (1, 3):
        row = [f"{name} {t(M, C, C, b, r, nset):6.1f} us" for name, b, r in (("bias+res", 1, 1), ("bias", 1, 0), ("none", 0, 0), ("res", 0, 1))]
        print(f"M={M} N=K={C} buffers x{nset}: " + " | ".join(row), flush=True)
