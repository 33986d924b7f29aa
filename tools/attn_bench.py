"""time ls_attention at the UNet's spatial self-attention shapes (CUDA events, L2-cold rotation of buffers)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentsync_b200 import _lib as L

dev = "cuda:0"
shapes = [(32, 8, 40, 1024), (32, 8, 80, 256), (64, 8, 40, 4096), (64, 8, 80, 1024), (64, 8, 160, 256)]
for batch, heads, d, S in shapes[: int(os.environ.get("NSHAPES", "2"))]:
    C = heads * d
    rows = batch * S
    nbuf = 6
    qkvs = [(torch.randn(rows, 3 * C, device=dev) * 1.5).half() for _ in range(nbuf)]
    outs = [torch.empty(rows, C, dtype=torch.float16, device=dev) for _ in range(nbuf)]
    def run(i):
        qkv, out = qkvs[i % nbuf], outs[i % nbuf]
        L.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, 3 * C, 3 * C, 3 * C, C, batch, heads, d, S, S)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 4.0 * batch * heads * S * S * d
    x = qkvs[0].float().reshape(batch, S, 3, heads, d).permute(2, 0, 3, 1, 4)[:, :2]
    ref = torch.softmax(x[0] @ x[1].transpose(-1, -2) * d ** -0.5, -1) @ x[2]
    run(0)
    got = outs[0].float().reshape(batch, S, heads, d).permute(0, 2, 1, 3)[:2]
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"LS_ATTN_TC={os.environ.get('LS_ATTN_TC','1')} batch={batch} heads={heads} d={d} S={S}: {us:8.1f} us  "
          f"{fl/us/1e6:7.1f} TFLOP/s  rel-L2 {err:.2e}", flush=True)
