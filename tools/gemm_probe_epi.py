import sys, os, math
os.environ.setdefault("LS_SO_NAME", "_C_gprobe.so")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev="cuda"
for (M,K,N,bn,res) in ((32768,64,320,128,0),(32768,320,320,128,1),(32768,64,2560,256,0)):
    a=torch.randn(M,K,device=dev).half(); w=(torch.randn(N,K,device=dev)/math.sqrt(K)).half()
    b=torch.randn(N,device=dev); r=torch.randn(M,N,device=dev).half() if res else None
    o=torch.empty(M,N,dtype=torch.float16,device=dev)
    for i in range(2):
        if i == 1: print(f"--- M={M} K={K} N={N} bn={bn} residual={res}", flush=True)
        L.gemm([L.Seg(a,K,K,1)],1,1,M,w,N,o,N,bias=b,residual=r,ldr=N,tile_n=bn,cta_pair=1)
        torch.cuda.synchronize()
