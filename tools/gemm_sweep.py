"""Times ls_gemm for the UNet's GEMM shapes under every (cta_pair, tile_n) choice.  Each config is captured as a CUDA
graph of REPS back-to-back launches (distinct output buffers rotate so the L2 state resembles the real forward)."""
import sys, os, math, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L

REPS = 20
dev = "cuda"
# (name, nimg, H, W, Cin, taps, N, geglu, residual)
SHAPES = [
    ("lin_L0_320", 1, 1, 32768, 320, 1, 320, 0, 1),
    ("qkv_L0", 1, 1, 32768, 320, 1, 960, 0, 0),
    ("geglu_L0", 1, 1, 32768, 320, 1, 2560, 1, 0),
    ("ff2_L0", 1, 1, 32768, 1280, 1, 320, 0, 1),
    ("lin_L1_640", 1, 1, 8192, 640, 1, 640, 0, 1),
    ("qkv_L1", 1, 1, 8192, 640, 1, 1920, 0, 0),
    ("geglu_L1", 1, 1, 8192, 640, 1, 5120, 1, 0),
    ("ff2_L1", 1, 1, 8192, 2560, 1, 640, 0, 1),
    ("lin_L2_1280", 1, 1, 2048, 1280, 1, 1280, 0, 1),
    ("geglu_L2", 1, 1, 2048, 1280, 1, 10240, 1, 0),
    ("ff2_L2", 1, 1, 2048, 5120, 1, 1280, 0, 1),
    ("lin_L3_1280", 1, 1, 512, 1280, 1, 1280, 0, 1),
    ("conv_L0_320", 32, 32, 32, 320, 9, 320, 0, 0),
    ("conv_L0_640_320", 32, 32, 32, 640, 9, 320, 0, 0),
    ("conv_L1_640", 32, 16, 16, 640, 9, 640, 0, 0),
    ("conv_L2_1280", 32, 8, 8, 1280, 9, 1280, 0, 0),
    ("conv_L3_1280", 32, 4, 4, 1280, 9, 1280, 0, 0),
    ("conv_L3_2560", 32, 4, 4, 2560, 9, 1280, 0, 0),
]
only = sys.argv[1:] 
out = {}
for name, nimg, H, W, Cin, taps, N, geglu, resid in SHAPES:
    if only and not any(o in name for o in only):
        continue
    M = nimg * H * W
    K = Cin * taps
    a = torch.randn(M, Cin, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    bias = torch.randn(N, device=dev)
    nout = N // 2 if geglu else N
    outs = [torch.empty(M, nout, dtype=torch.float16, device=dev) for _ in range(4)]
    res = torch.randn(M, nout, device=dev).half() if resid else None
    flops = 2.0 * M * N * K
    best = None
    for ctas in (1, 2):
        for bn in range(64 if geglu else 32, 257, 64 if geglu else 32):
            if geglu and N % bn:
                continue
            if bn - 32 >= N:
                continue
            def launch(i):
                L.gemm([L.Seg(a, Cin, Cin, taps)], nimg, H, W, w, N, outs[i % 4], nout, bias=bias, residual=res, ldr=nout,
                       flags=L.EPI_GEGLU if geglu else 0, tile_n=bn, cta_pair=ctas)
            try:
                launch(0); launch(1)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(REPS):
                        launch(i)
                g.replay(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1000 / (2 * REPS)
            except Exception as ex:
                print(name, ctas, bn, "FAILED", str(ex)[:100]); continue
            out.setdefault(name, []).append((ctas, bn, us))
            if best is None or us < best[2]:
                best = (ctas, bn, us)
    # what does auto pick?
    def launch_auto(i):
        L.gemm([L.Seg(a, Cin, Cin, taps)], nimg, H, W, w, N, outs[i % 4], nout, bias=bias, residual=res, ldr=nout,
               flags=L.EPI_GEGLU if geglu else 0)
    launch_auto(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS):
            launch_auto(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us_auto = e0.elapsed_time(e1) * 1000 / (2 * REPS)
    rows = sorted(out[name], key=lambda r: r[2])
    print(f"{name:18s} M={M} N={N} K={K}: best ctas={best[0]} bn={best[1]} {best[2]:.1f}us {flops/best[2]/1e6:.0f} TF/s | auto {us_auto:.1f}us | " +
          " ".join(f"{c}/{b}:{u:.1f}" for c, b, u in rows[:6]))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_sweep.json"), "w"))
