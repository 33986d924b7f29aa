import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentsync_b200 import _lib as L
dev = "cuda:0"
torch.manual_seed(0)
d, heads = int(os.environ.get("D", "40")), 8
C = heads * d
shapes = [(1, 128, 64), (1, 128, 128), (1, 128, 192), (1, 128, 256), (1, 128, 320), (1, 128, 512), (2, 256, 1024), (32, 1024, 1024)]
if os.environ.get("RAGGED"):
    shapes = [(2, 300, 200), (3, 200, 130), (2, 130, 65), (2, 256, 256), (4, 256, 256)]
for batch, sq, skv in shapes:
    q = torch.randn(batch * sq, C, device=dev).half() * float(os.environ.get("QS", "1.0"))
    k = torch.randn(batch * skv, C, device=dev).half()
    v = torch.randn(batch * skv, C, device=dev).half()
    out = torch.zeros(batch * sq, C, dtype=torch.float16, device=dev)
    try:
        L.attention(q, k, v, out, C, C, C, C, batch, heads, d, sq, skv)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"batch={batch} sq={sq} skv={skv}: FAILED {str(e)[:80]}")
        break
    qh = (q.float() * d ** -0.5).reshape(batch, sq, heads, d).transpose(1, 2)
    kh = k.float().reshape(batch, skv, heads, d).transpose(1, 2)
    vh = v.float().reshape(batch, skv, heads, d).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(batch * sq, C)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    rows = (out.float() - ref).norm(dim=1) / ref.norm(dim=1)
    print(f"batch={batch} sq={sq} skv={skv}: rel-L2 {err:.2e}; bad rows {(rows > 1e-2).sum().item()} of {rows.numel()}; first bad {torch.nonzero(rows > 1e-2)[:4].flatten().tolist()}", flush=True)
