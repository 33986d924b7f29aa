"""ncu target for the GroupNorm-from-the-epilogue pair: the level-0 short-K linear (bias + residual) without and with
gn_partials_out, then the joint GroupNorm + SiLU from those partials, then the per-frame one.

  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'gn_parts_kernel|gemm_tc_kernel' \
      -o gpurun_out/r2s_gn_parts python tools/gn_parts_one.py
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import _lib as L  # noqa: E402

dev = "cuda"
M, K, N, unit = 32768, 320, 320, 10
a = torch.randn(M, K, device=dev).half()
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
bias = torch.randn(N, device=dev)
res = torch.randn(M, N, device=dev).half()
out = torch.empty(M, N, dtype=torch.float16, device=dev)
parts = torch.empty(M // 128, N // unit, 2, dtype=torch.float32, device=dev)
gamma, beta = torch.randn(N, device=dev), torch.randn(N, device=dev)
y = torch.empty_like(out)


def run():
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out, N, bias=bias, residual=res, ldr=N, tile_n=160)
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out, N, bias=bias, residual=res, ldr=N, gn_partials_out=parts, gn_unit=unit)
    L.groupnorm_parts(out, N, parts, None, 0, None, M, M // 2, 32, unit, gamma, beta, 1e-5, True, y)
    L.groupnorm_parts(out, N, parts, None, 0, None, M, 1024, 32, unit, gamma, beta, 1e-6, False, y)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
