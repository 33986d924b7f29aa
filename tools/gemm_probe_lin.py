import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev="cuda"
M,K,N=32768,320,320
a=torch.randn(M,K,device=dev).half(); w=(torch.randn(N,K,device=dev)/math.sqrt(K)).half()
b=torch.randn(N,device=dev); r=torch.randn(M,N,device=dev).half(); o=torch.empty(M,N,dtype=torch.float16,device=dev)
for name,bb,rr in (("none",None,None),("bias+res",b,r)):
    print(f"--- {name}", flush=True)
    for i in range(2):
        L.gemm([L.Seg(a,K,K,1)],1,1,M,w,N,o,N,bias=bb,residual=rr,ldr=N)
        torch.cuda.synchronize()
