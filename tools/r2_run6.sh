#!/bin/bash
# round 2, GPU call 6: one vs two k-blocks per pipeline stage, per shape
mkdir -p gpurun_out
LS_GEMM_KBS=2 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -5 > gpurun_out/r2e_ops_gemm_kbs2.log
cat gpurun_out/r2e_ops_gemm_kbs2.log
LS_GEMM_KBS=1 timeout 900 python tools/gemm_shapes.py --bns 64,128,160,256 > gpurun_out/r2e_shapes_kbs1.txt 2>&1
LS_GEMM_KBS=2 timeout 900 python tools/gemm_shapes.py --bns 64,128,160,256 > gpurun_out/r2e_shapes_kbs2.txt 2>&1
LS_GEMM_KBS=2 timeout 600 python bench.py --steps 3 --no-extras > gpurun_out/r2e_bench_kbs2.json 2> gpurun_out/r2e_bench_kbs2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_bench_kbs2.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("kbs2: fps", round(d["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3)
PY
