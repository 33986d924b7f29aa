"""time ls_groupnorm at the UNet's shapes inside a CUDA graph (rotating buffers)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentsync_b200 import _lib as L
dev = "cuda"
REPS = 20
shapes = [(32768, 320, 16384, 1), (32768, 320, 1024, 0), (8192, 640, 4096, 1), (8192, 640, 256, 0), (2048, 1280, 1024, 1),
          (2048, 1280, 64, 0), (512, 1280, 256, 1), (512, 1280, 16, 0), (512, 2560, 256, 1), (16384, 512, 1024, 1),
          (32768, 960, 16384, 1), (1048576, 128, 65536, 1)]
lib = L.lib()
for rows, C, rpi, silu in shapes:
    xs = [torch.randn(rows, C, device=dev).half() for _ in range(3)]
    ys = [torch.empty(rows, C, dtype=torch.float16, device=dev) for _ in range(3)]
    g = torch.rand(C, device=dev) + 0.5
    b = torch.randn(C, device=dev) * 0.1
    stats = torch.zeros((rows // rpi) * 32 * 2, device=dev)
    def run(i):
        L._check(lib.ls_groupnorm(xs[i % 3].data_ptr(), C, None, 0, rows, rpi, 32, g.data_ptr(), b.data_ptr(), 1e-5, silu,
                                  stats.data_ptr(), ys[i % 3].data_ptr(), torch.cuda.current_stream().cuda_stream), "gn")
    run(0); run(1); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(REPS):
            run(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (2 * REPS)
    mb = rows * C * 2 / 1e6
    x = xs[0].float().view(rows // rpi, rpi, 32, C // 32)
    m = x.mean(dim=(1, 3), keepdim=True); v = x.var(dim=(1, 3), unbiased=False, keepdim=True)
    ref = ((x - m) / torch.sqrt(v + 1e-5)).view(rows, C) * g + b
    if silu: ref = torch.nn.functional.silu(ref)
    run(0); torch.cuda.synchronize()
    err = ((ys[0].float() - ref).norm() / ref.norm()).item()
    print(f"rows={rows} C={C} rows_per_inst={rpi} silu={silu}: {us:7.1f} us  ({2*mb/us:5.2f} TB/s read+write of {mb:.1f} MB)  rel-L2 {err:.1e}", flush=True)
