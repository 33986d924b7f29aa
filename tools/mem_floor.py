"""memory-system floors for comparison with the GEMM epilogue: write-only, copy, read-only at the sizes of the UNet's
activation tensors (rotating 4 buffers like tools/gemm_small_k.py)"""
import torch
dev = "cuda"
REPS = 20
def timeit(fn):
    fn(0); fn(1); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * REPS)
for mb in (5.2, 21, 63, 168):
    n = int(mb * 1e6 / 2)
    bufs = [torch.empty(n, dtype=torch.float16, device=dev) for _ in range(4)]
    src = torch.randn(n, device=dev).half()
    t_fill = timeit(lambda i: bufs[i % 4].fill_(1.0))
    t_copy = timeit(lambda i: bufs[i % 4].copy_(src))
    t_add = timeit(lambda i: torch.add(src, bufs[(i + 1) % 4], out=bufs[i % 4]))
    print(f"{mb:6.1f} MB: fill {t_fill:6.1f} us ({mb/t_fill:5.2f} TB/s written)  copy {t_copy:6.1f} us ({2*mb/t_copy:5.2f} TB/s r+w)  "
          f"add(2 reads, 1 write) {t_add:6.1f} us ({3*mb/t_add:5.2f} TB/s)", flush=True)
