"""per-launch floor of each kernel family inside a CUDA graph (tiny problem sizes)"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
def timeit(fn, reps=50):
    fn(); fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * reps)
for (M, K, N) in ((128, 64, 32), (512, 1280, 1280), (2048, 1280, 1280), (32768, 320, 320), (8192, 640, 640)):
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); o = torch.empty(M, N, device=dev, dtype=torch.float16)
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).half()
    print(f"gemm M={M} K={K} N={N}: {timeit(lambda: L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, o, N, bias=b, residual=r, ldr=N)):.2f} us")
x = torch.randn(512, 1280, device=dev).half(); y = torch.empty_like(x); g_ = torch.ones(1280, device=dev); b_ = torch.zeros(1280, device=dev)
print(f"layernorm 512x1280: {timeit(lambda: L.layernorm(x, 512, 1280, g_, b_, 1e-5, y)):.2f} us")
x = torch.randn(32768, 320, device=dev).half(); y = torch.empty_like(x); g_ = torch.ones(320, device=dev); b_ = torch.zeros(320, device=dev)
print(f"layernorm 32768x320: {timeit(lambda: L.layernorm(x, 32768, 320, g_, b_, 1e-5, y)):.2f} us")
st = torch.empty(32 * 64, device=dev)
print(f"groupnorm 32768x320 (stats+apply): {timeit(lambda: L.groupnorm(x, 320, None, 0, 32768, 16384, 32, g_, b_, 1e-5, True, y, st)):.2f} us")
x2 = torch.randn(512, 1280, device=dev).half(); y2 = torch.empty_like(x2); g2 = torch.ones(1280, device=dev); b2 = torch.zeros(1280, device=dev)
print(f"groupnorm 512x1280 (stats+apply): {timeit(lambda: L.groupnorm(x2, 1280, None, 0, 512, 256, 32, g2, b2, 1e-5, True, y2, st)):.2f} us")
q = torch.randn(32768, 960, device=dev).half(); o = torch.empty(32768, 320, device=dev, dtype=torch.float16)
print(f"attention spatial S=1024 d=40: {timeit(lambda: L.attention(q[:, :320], q[:, 320:640], q[:, 640:], o, 960, 960, 960, 320, 32, 8, 40, 1024, 1024), 10):.2f} us")
print(f"attention temporal L0: {timeit(lambda: L.attention(q[:, :320], q[:, 320:640], q[:, 640:], o, 960, 960, 960, 320, 2048, 8, 40, 16, 16, q_addr=(1024, 16384, 1, 1024), kv_addr=(1024, 16384, 1, 1024))):.2f} us")
