"""Two launches for an ncu source-level capture: the level-0 GEGLU projection (M=32768, N=2560, K=320) and the
level-0 short-K linear with residual (M=32768, N=320, K=320).  Each is run 4 times; capture with -s 4 -c 1 / -s 7 -c 1."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
M, K = 32768, 320
a = torch.randn(M, K, device=dev).half()
# GEGLU
N = 2560
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
b = torch.randn(N, device=dev)
o = torch.empty(M, N // 2, dtype=torch.float16, device=dev)
for i in range(4):
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, o, N // 2, bias=b, flags=1, tile_n=256, cta_pair=1)
torch.cuda.synchronize()
# plain short-K linear with residual
N = 320
w2 = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
b2 = torch.randn(N, device=dev)
r = torch.randn(M, N, device=dev).half()
o2 = torch.empty(M, N, dtype=torch.float16, device=dev)
for i in range(4):
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w2, N, o2, N, bias=b2, residual=r, ldr=N, tile_n=160, cta_pair=1)
torch.cuda.synchronize()
print("ok")
