#!/bin/bash
# round 2, GPU call 3: coalesced residual epilogue + retuned tables: gemm tests, shape sweep subset, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -5 > gpurun_out/r2c_ops_gemm.log
cat gpurun_out/r2c_ops_gemm.log
timeout 600 python tools/gemm_shapes.py --only lin --bns 128,160 > gpurun_out/r2c_shapes_lin.txt 2>&1
timeout 600 python tools/gemm_shapes.py --only ffout --bns 160 >> gpurun_out/r2c_shapes_lin.txt 2>&1
grep -E "^lin|^ffout|ctas=0|best" gpurun_out/r2c_shapes_lin.txt
timeout 900 python bench.py --steps 5 --no-extras --profile-kernels > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), r["other_kinds_ms_in_graph"], "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3)
PY
grep "^vae   gemm\|^unet  gemm" gpurun_out/r2c_bench.err
