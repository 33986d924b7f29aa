"""one epilogue-bound launch for ncu: the level-0 short-K linear (M = 32768, N = K = 320, bias + residual), rotating buffers"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
M, N, K = int(os.environ.get("M", 32768)), int(os.environ.get("N", 320)), int(os.environ.get("K", 320))
a = [torch.randn(M, K, device=dev).half() for _ in range(3)]
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
b = torch.randn(N, device=dev)
r = torch.randn(M, N, device=dev).half()
o = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(3)]
for i in range(6):
    L.gemm([L.Seg(a[i % 3], K, K, 1)], 1, 1, M, w, N, o[i % 3], N, bias=b, residual=r, ldr=N)
torch.cuda.synchronize()
print("ok")
