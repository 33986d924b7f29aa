"""one main-loop-bound launch for ncu: the level-0 3x3 convolution (32 x 32 x 32 images, 320 -> 320 channels, K = 2880)"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
nimg, H, W, Cin, N = 32, 32, 32, 320, 320
M, K = nimg * H * W, 9 * Cin
a = torch.randn(M, Cin, device=dev).half()
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
b = torch.randn(N, device=dev)
o = torch.empty(M, N, dtype=torch.float16, device=dev)
for i in range(4):
    L.gemm([L.Seg(a, Cin, Cin, 9)], nimg, H, W, w, N, o, N, bias=b)
torch.cuda.synchronize()
print("ok")
