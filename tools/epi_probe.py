"""in-graph time of the UNet's epilogue-bound GEMM shapes (auto tile selection), for before/after comparisons"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 20
def t(M, N, K, resid, bias, geglu=0, bn=0):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    b = torch.randn(N, device=dev) if bias else None
    nout = N // 2 if geglu else N
    outs = [torch.empty(M, nout, dtype=torch.float16, device=dev) for _ in range(4)]
    res = torch.randn(M, nout, device=dev).half() if resid else None
    def launch(i):
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, outs[i % 4], nout, bias=b, residual=res, ldr=nout,
               flags=L.EPI_GEGLU if geglu else 0, tile_n=bn if bn else (256 if geglu else 0))
    launch(0); launch(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS):
            launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * REPS)
shapes = [("geglu_L0", 32768, 2560, 320, 0, 1, 1), ("geglu_L1", 8192, 5120, 640, 0, 1, 1), ("geglu_L2", 2048, 10240, 1280, 0, 1, 1),
          ("lin_L0", 32768, 320, 320, 1, 1, 0), ("lin_L1", 8192, 640, 640, 1, 1, 0), ("lin_L2", 2048, 1280, 1280, 1, 1, 0),
          ("lin_L3", 512, 1280, 1280, 1, 1, 0),
          ("qkv_L0", 32768, 960, 320, 0, 0, 0), ("qkv_L1", 8192, 1920, 640, 0, 0, 0), ("ff2_L0", 32768, 320, 1280, 1, 1, 0),
          ("ff2_L1", 8192, 640, 2560, 1, 1, 0)]
tot = 0.0
for name, M, N, K, r, b, g in shapes:
    us = t(M, N, K, r, b, g)
    fl = 2.0 * M * N * K
    print(f"{name:9s} M={M:6d} N={N:6d} K={K:5d}: {us:7.1f} us  {fl / us * 1e-6:7.1f} TF/s", flush=True)
