"""where does the time of the K=320 linears go?  same M, N with / without residual and bias, and with K from 64 up"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 20
def t(M, N, K, bn, resid, bias, geglu=0, ctas=1):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    b = torch.randn(N, device=dev) if bias else None
    nout = N // 2 if geglu else N
    outs = [torch.empty(M, nout, dtype=torch.float16, device=dev) for _ in range(4)]
    res = torch.randn(M, nout, device=dev).half() if resid else None
    def launch(i):
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, outs[i % 4], nout, bias=b, residual=res, ldr=nout,
               flags=L.EPI_GEGLU if geglu else 0, tile_n=bn, cta_pair=ctas)
    launch(0); launch(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS):
            launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * REPS)
M = 32768
for N, bn in ((320, 128), (320, 64), (320, 192), (960, 192), (960, 256)):
    for K in (64, 128, 320, 640, 1280):
        row = []
        for resid, bias in ((0, 0), (0, 1), (1, 1)):
            row.append(t(M, N, K, bn, resid, bias))
        print(f"M={M} N={N} bn={bn} K={K:5d}: none {row[0]:6.1f}  bias {row[1]:6.1f}  bias+res {row[2]:6.1f} us", flush=True)
for K in (64, 320):
    for bn in (128, 256):
        print(f"geglu M={M} N=2560 bn={bn} K={K}: {t(M, 2560, K, bn, 0, 1, geglu=1):6.1f} us   plain N=2560: {t(M, 2560, K, bn, 0, 1):6.1f} us", flush=True)
