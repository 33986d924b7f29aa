"""level-3 (M = 512) GEMM shapes with COLD weights (8 weight copies rotate: 236 MB > L2 for the conv), auto tile /
split selection; run once per LS_GEMM_SPLITK value (the env var is read once per process)"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 16
SHAPES = [("lin_L3", 1, 1, 512, 1280, 1, 1280, 1), ("qkv_L3", 1, 1, 512, 1280, 1, 3840, 0), ("ff2_L3", 1, 1, 512, 5120, 1, 1280, 1),
          ("conv_L3_1280", 32, 4, 4, 1280, 9, 1280, 0), ("conv_L3_2560", 32, 4, 4, 2560, 9, 1280, 0),
          ("conv_L2_1280", 32, 8, 8, 1280, 9, 1280, 0), ("lin_L2", 1, 1, 2048, 1280, 1, 1280, 1)]
bn = int(os.environ.get("BN", "0"))
for name, nimg, H, W, Cin, taps, N, resid in SHAPES:
    M = nimg * H * W; K = Cin * taps
    a = torch.randn(M, Cin, device=dev).half()
    ws = [(torch.randn(N, K, device=dev) / math.sqrt(K)).half() for _ in range(8)]
    bias = torch.randn(N, device=dev)
    outs = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(4)]
    res = torch.randn(M, N, device=dev).half() if resid else None
    def launch(i):
        L.gemm([L.Seg(a, Cin, Cin, taps)], nimg, H, W, ws[i % 8], N, outs[i % 4], N, bias=bias, residual=res, ldr=N, tile_n=bn)
    launch(0); launch(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS): launch(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (2 * REPS)
    wb = N * K * 2 / 1e6
    print(f"{name:13s} M={M:5d} N={N:5d} K={K:6d}: {us:7.1f} us  {2.0*M*N*K/us*1e-6:7.1f} TF/s  weights {wb:5.1f} MB -> {wb/us*1e-3:5.2f} TB/s", flush=True)
