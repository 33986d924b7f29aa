import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev="cuda"
def run(name,nimg,H,W,Cin,taps,N,bn,ctas):
    M=nimg*H*W;K=Cin*taps
    a=torch.randn(M,Cin,device=dev).half(); w=(torch.randn(N,K,device=dev)/math.sqrt(K)).half()
    outs=[torch.empty(M,N,dtype=torch.float16,device=dev) for _ in range(4)]
    def launch(i): L.gemm([L.Seg(a,Cin,Cin,taps)],nimg,H,W,w,N,outs[i%4],N,tile_n=bn,cta_pair=ctas)
    launch(0);launch(1);torch.cuda.synchronize()
    if os.environ.get('LS_PROBE'): return
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(20): launch(i)
    g.replay();torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record();g.replay();g.replay();e1.record();torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*1000/40
    print(f"dbg={os.environ.get('LS_GEMM_DBG','0')} {name} bn={bn} ctas={ctas}: {us:.1f} us")
for bn,ct in ((256,2),(160,2)):
    run("conv_L0_320",32,32,32,320,9,320,bn,ct)
    run("ff2_L1",1,1,8192,2560,1,640,bn,ct)
