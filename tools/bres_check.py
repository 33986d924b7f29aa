"""weight-resident GEMM mode: correctness against an fp32 reference on the shapes that take it"""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import _lib as L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(3)
for (M, N, K, res, geglu) in ((32768, 320, 320, True, False), (32768, 960, 320, False, False), (16384, 320, 320, True, False),
                              (32768 + 64, 320, 320, True, False), (32768, 2560, 320, False, True), (40000, 320, 64, True, False)):
    a = torch.randn(M, K, device=dev, generator=g).half()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K))
    b = torch.randn(N, device=dev, generator=g)
    n_out = N // 2 if geglu else N
    r = torch.randn(M, n_out, device=dev, generator=g).half() if res else None
    o = torch.empty(M, n_out, dtype=torch.float16, device=dev)
    if geglu:
        wp, bp = L.pack_geglu(w, b, 256)
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, wp.half().contiguous(), N, o, n_out, bias=bp, flags=L.EPI_GEGLU, tile_n=256)
        h = a.float() @ w.half().float().t() + b
        want = h[:, :n_out] * torch.nn.functional.gelu(h[:, n_out:])
    else:
        L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w.half().contiguous(), N, o, n_out, bias=b, residual=r, ldr=n_out)
        want = a.float() @ w.half().float().t() + b + (r.float() if res else 0)
    err = ((o.float() - want).norm() / want.norm()).item()
    print(f"M={M} N={N} K={K} res={res} geglu={geglu}: rel-L2 {err:.2e}", "OK" if err < 2e-3 else "FAIL", flush=True)
