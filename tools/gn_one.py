import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentsync_b200 import _lib as L
dev = "cuda"
rows, C, rpi, silu = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (32768, 320, 1024, 0)
lib = L.lib()
x = torch.randn(rows, C, device=dev).half(); y = torch.empty_like(x)
g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
stats = torch.zeros((rows // rpi) * 64, device=dev)
for i in range(6):
    L._check(lib.ls_groupnorm(x.data_ptr(), C, None, 0, rows, rpi, 32, g.data_ptr(), b.data_ptr(), 1e-5, silu,
                              stats.data_ptr(), y.data_ptr(), torch.cuda.current_stream().cuda_stream), "gn")
torch.cuda.synchronize()
print("ok")
