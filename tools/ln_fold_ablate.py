"""Where the folded-LayerNorm epilogues spend their time: in-graph us per launch of consumer / producer variants on the
level-0 / level-1 shapes (A/B tool; L2-warm operands).

    python tools/ln_fold_ablate.py > gpurun_out/ln_fold_ablate.txt
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import _lib as L  # noqa: E402

dev = "cuda"
REPS = 8


def timed(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / (2 * REPS)


def consumer(M, C, N, geglu):
    x = torch.randn(M, C, device=dev).half()
    w = (torch.randn(N, C, device=dev) / math.sqrt(C)).half()
    bias = torch.randn(N, device=dev)
    cs = torch.randn(N, device=dev)
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, dtype=torch.float16, device=dev)
    flags = L.EPI_GEGLU if geglu else 0
    res = []
    for tn in ((256,) if geglu else (0, 160, 192, 256)):
        row = [f"tile_n={tn:3d}"]
        for name, kw in (("plain", {}), ("plain+bias", {"bias": bias}),
                         ("ln 1 part", {"bias": bias, "col_sum": cs, "row_partials_in": torch.rand(1, M, 2, device=dev) + 1}),
                         ("ln 6 parts", {"bias": bias, "col_sum": cs, "row_partials_in": torch.rand(6, M, 2, device=dev) + 1}),
                         ("ln 12 parts", {"bias": bias, "col_sum": cs, "row_partials_in": torch.rand(12, M, 2, device=dev) + 1})):
            us = timed(lambda: L.gemm([L.Seg(x, C, C, 1)], 1, 1, M, w, N, out, n_out, flags=flags, tile_n=tn, **kw))
            row.append(f"{name} {us:6.1f}")
        res.append(" | ".join(row))
    print(f"consumer M={M} C={C} N={N} geglu={int(geglu)}")
    for r in res:
        print("   ", r, flush=True)


def producer(M, C, K):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(C, K, device=dev) / math.sqrt(K)).half()
    bias = torch.randn(C, device=dev)
    r = torch.randn(M, C, device=dev).half()
    out = torch.empty(M, C, dtype=torch.float16, device=dev)
    parts = torch.empty(3 * ((C + 159) // 160), M, 2, device=dev)
    row = []
    for name, kw in (("bias", {}), ("bias+stats", {"row_partials_out": parts}), ("bias+res", {"residual": r, "ldr": C}),
                     ("bias+res+stats", {"residual": r, "ldr": C, "row_partials_out": parts})):
        us = timed(lambda: L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, C, out, C, bias=bias, tile_n=160, **kw))
        row.append(f"{name} {us:6.1f}")
    print(f"producer M={M} N={C} K={K}: " + " | ".join(row), flush=True)


consumer(32768, 320, 960, False)
consumer(32768, 320, 2560, True)
consumer(8192, 640, 1920, False)
consumer(8192, 640, 5120, True)
producer(32768, 320, 320)
producer(32768, 320, 1280)
producer(8192, 640, 640)
producer(2048, 1280, 1280)
