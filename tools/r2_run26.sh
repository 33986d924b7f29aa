#!/bin/bash
mkdir -p gpurun_out
echo "== check BRES=1"; LS_GEMM_BRES=1 timeout 300 python tools/bres_check.py
LS_GEMM_BRES=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -4
echo "BRES=0"; LS_GEMM_BRES=0 python tools/epi_ablate.py | grep "x1"
echo "BRES=1"; LS_GEMM_BRES=1 python tools/epi_ablate.py | grep "x1"
for v in 0 1; do for k in lin qkv geglu; do LS_GEMM_BRES=$v python tools/gemm_shapes.py --bns 160 --only $k 2>&1 | grep -E "ctas=0|^[a-z]" | paste - - | awk -v v=$v '{print "BRES=" v, $0}' | cut -c1-150; done; done
LS_GEMM_BRES=1 timeout 600 python bench.py --steps 3 --no-extras > gpurun_out/r2o_bench_bres.json 2> gpurun_out/r2o_bench_bres.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench_bres.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("BRES=1: fps", round(d["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3)
PY
