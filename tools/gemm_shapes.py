"""Sweep of the UNet's real GEMM / conv shapes over (CTA pairing, tile width): in-graph time per launch, TFLOP/s and the
max |difference| against the library's default choice (correctness of every forced configuration).

    LS_SO_NAME=_C_lane0.so python tools/gemm_shapes.py > gpurun_out/shapes_lane0.txt
    python tools/gemm_shapes.py [--quick] [--bns 128,160,256]

Output feeds the host cost model in csrc/gemm_tc.cu (which (ctas, BN) to pick per shape)."""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import _lib as L  # noqa: E402

dev = "cuda"
REPS = 6

# (name, nimg, H, W, [(channels, taps)...], N, flags, residual)
SHAPES = [
    ("lin L0 C=320", 1, 1, 32768, [(320, 1)], 320, 0, True),
    ("lin L1 C=640", 1, 1, 8192, [(640, 1)], 640, 0, True),
    ("lin L2 C=1280", 1, 1, 2048, [(1280, 1)], 1280, 0, True),
    ("lin L3 C=1280", 1, 1, 512, [(1280, 1)], 1280, 0, True),
    ("geglu L0", 1, 1, 32768, [(320, 1)], 2560, 1, False),
    ("geglu L1", 1, 1, 8192, [(640, 1)], 5120, 1, False),
    ("geglu L2", 1, 1, 2048, [(1280, 1)], 10240, 1, False),
    ("qkv L0", 1, 1, 32768, [(320, 1)], 960, 0, False),
    ("qkv L1", 1, 1, 8192, [(640, 1)], 1920, 0, False),
    ("qkv L2", 1, 1, 2048, [(1280, 1)], 3840, 0, False),
    ("ffout L0", 1, 1, 32768, [(1280, 1)], 320, 0, True),
    ("ffout L1", 1, 1, 8192, [(2560, 1)], 640, 0, True),
    ("ffout L2", 1, 1, 2048, [(5120, 1)], 1280, 0, True),
    ("conv L0 320->320", 32, 32, 32, [(320, 9)], 320, 0, True),
    ("conv L0 640->320", 32, 32, 32, [(640, 9)], 320, 0, False),
    ("conv L0 320->320+sc", 32, 32, 32, [(320, 9), (320, 1), (320, 1)], 320, 0, False),
    ("conv L0 640->640 up", 32, 32, 32, [(640, 9)], 640, 0, False),
    ("conv L1 640->640", 32, 16, 16, [(640, 9)], 640, 0, True),
    ("conv L1 1280->640", 32, 16, 16, [(1280, 9)], 640, 0, False),
    ("conv L1 1280->1280 up", 32, 16, 16, [(1280, 9)], 1280, 0, False),
    ("conv L2 1280->1280", 32, 8, 8, [(1280, 9)], 1280, 0, True),
    ("conv L2 2560->1280", 32, 8, 8, [(2560, 9)], 1280, 0, False),
    ("conv L3 1280->1280", 32, 4, 4, [(1280, 9)], 1280, 0, True),
    ("conv L3 2560->1280", 32, 4, 4, [(2560, 9)], 1280, 0, False),
    ("vae 512 32x32", 16, 32, 32, [(512, 9)], 512, 0, True),
    ("vae 512 64x64", 16, 64, 64, [(512, 9)], 512, 0, True),
    ("vae 512->256 128x128", 16, 128, 128, [(512, 9)], 256, 0, False),
    ("vae 256 128x128", 16, 128, 128, [(256, 9)], 256, 0, True),
    ("vae 256->128 256x256", 16, 256, 256, [(256, 9)], 128, 0, False),
    ("vae 128 256x256", 16, 256, 256, [(128, 9)], 128, 0, True),
]


def bench_shape(sh, configs):
    name, nimg, H, W, segs, N, flags, use_res = sh
    M = nimg * H * W
    K = sum(c * t for c, t in segs)
    nset = 3
    a_sets = [[torch.randn(M, c, device=dev).half() for c, _ in segs] for _ in range(nset)]
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    bias = torch.randn(N, device=dev)
    n_out = N // 2 if flags & 1 else N
    res = torch.randn(M, n_out, device=dev).half() if use_res else None
    outs = [torch.empty(M, n_out, dtype=torch.float16, device=dev) for _ in range(nset)]

    def launch(i, ctas, bn):
        s = a_sets[i % nset]
        L.gemm([L.Seg(t, c, c, tp) for t, (c, tp) in zip(s, segs)], nimg, H, W, w, N, outs[i % nset], n_out, bias=bias,
               residual=res, ldr=n_out, flags=flags, tile_n=bn, cta_pair=ctas)

    launch(0, 0, 256 if flags & 1 else 0)
    torch.cuda.synchronize()
    ref = outs[0].float().clone()
    flops = 2.0 * M * N * K
    rows = []
    for ctas, bn in configs:
        if flags & 1 and bn == 0:
            bn = 256
        try:
            launch(0, ctas, bn)
            torch.cuda.synchronize()
            err = (outs[0].float() - ref).abs().max().item()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(REPS):
                    launch(i, ctas, bn)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1000 / (2 * REPS)
            rows.append((ctas, bn, us, flops / us / 1e6, err))
        except RuntimeError as e:  # configuration rejected by the host checks
            rows.append((ctas, bn, float("nan"), 0.0, str(e).splitlines()[-1][:60]))
    return M, N, K, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--bns", default="64,96,128,160,192,224,256")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    bns = [int(b) for b in args.bns.split(",")]
    configs = [(0, 0)] + [(c, b) for c in (1, 2) for b in bns]
    shapes = [s for s in SHAPES if args.only in s[0]]
    if args.quick:
        shapes = shapes[::3]
    print(f"# so={os.environ.get('LS_SO_NAME', '_C.so')}  (ctas=0 / bn=0: the library's own choice)")
    for sh in shapes:
        M, N, K, rows = bench_shape(sh, configs)
        print(f"{sh[0]:24s} M={M} N={N} K={K} flags={sh[6]}")
        best = min((r for r in rows if r[2] == r[2]), key=lambda r: r[2])
        for ctas, bn, us, tf, err in rows:
            mark = " <== best" if (ctas, bn) == (best[0], best[1]) else ""
            e = f"{err:.2e}" if isinstance(err, float) else err
            print(f"    ctas={ctas} bn={bn:3d}: {us:8.1f} us {tf:7.1f} TF/s  maxdiff {e}{mark}", flush=True)


if __name__ == "__main__":
    main()
