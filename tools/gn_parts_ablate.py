"""GroupNorm statistics from the producing GEMM's epilogue: what the producer pays and what the norm saves, in-graph us
per launch on the UNet's shapes (A/B tool; operands L2-warm, i.e. the state a norm finds behind its producer).

    python tools/gn_parts_ablate.py > gpurun_out/gn_parts_ablate.txt
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


def producer(M, K, N, conv, side=32, unit=10, res=True):
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).half() if res else None
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    parts = torch.empty(M // 128, N // unit, 2, dtype=torch.float32, device=dev)
    if conv:
        nimg = M // (side * side)
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, 9 * K, device=dev) / math.sqrt(9 * K)).half()
        segs, geo = [L.Seg(a, K, K, 9)], (nimg, side, side)
    else:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
        segs, geo = [L.Seg(a, K, K, 1)], (1, 1, M)
    kw = dict(bias=bias, residual=r, ldr=N)
    t0 = timed(lambda: L.gemm(segs, *geo, w, N, out, N, **kw))
    t160 = timed(lambda: L.gemm(segs, *geo, w, N, out, N, tile_n=160 if unit == 10 else 0, **kw))
    t1 = timed(lambda: L.gemm(segs, *geo, w, N, out, N, gn_partials_out=parts, gn_unit=unit, **kw))
    print(f"producer M={M} K={K} N={N} conv={int(conv)} res={int(res)}: auto {t0:6.1f} us | tile 160 {t160:6.1f} | "
          f"+ GroupNorm partials {t1:6.1f}", flush=True)


def consumer(rows, c1, c2, rpi, silu, unit=10):
    C = c1 + c2
    x1 = torch.randn(rows, c1, device=dev).half()
    x2 = torch.randn(rows, c2, device=dev).half() if c2 else None
    p1 = torch.rand(rows // 128, c1 // unit, 2, device=dev)
    p2 = torch.rand(rows // 128, c2 // unit, 2, device=dev) if c2 else None
    p1[..., 1] += 1e4
    if p2 is not None:
        p2[..., 1] += 1e4
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.empty(rows, C, dtype=torch.float16, device=dev)
    stats = torch.empty(rows // rpi * 64, dtype=torch.float32, device=dev)
    t0 = timed(lambda: L.groupnorm_fused(x1, c1, x2, c2, rows, rpi, 32, gamma, beta, 1e-5, silu, out, stats))
    t1 = timed(lambda: L.groupnorm_parts(x1, c1, p1, x2, c2, p2, rows, rpi, 32, unit, gamma, beta, 1e-5, silu, out))
    print(f"norm rows={rows} C={c1}+{c2} rows_per_inst={rpi} silu={int(silu)}: self-contained {t0:6.1f} us | from partials "
          f"{t1:6.1f} us   ({2 * rows * C * 2 / 1e6:.0f} MB read + written)", flush=True)


if __name__ == "__main__":
    producer(32768, 320, 320, False)
    producer(8192, 640, 640, False)
    producer(2048, 1280, 1280, False)
    producer(32768, 320, 320, True, 32)
    producer(32768, 640, 320, True, 32, res=False)
    producer(8192, 640, 640, True, 16)
    producer(2048, 1280, 1280, True, 8)
    for lvl, (hw, c) in enumerate(((1024, 320), (256, 640), (64, 1280))):
        rows = 32 * hw
        if hw % 128 == 0:
            consumer(rows, c, 0, hw, False)
        consumer(rows, c, 0, 16 * hw, True)
    consumer(32768, 640, 320, 16384, True)
    consumer(32768, 320, 320, 16384, True)
    consumer(8192, 1280, 640, 4096, True)
    consumer(2048, 1280, 1280, 1024, True)
