#!/bin/bash
# round 2, GPU call 7: slab_single staging (one more pipeline stage for main-loop-bound launches)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -5 > gpurun_out/r2f_ops_gemm.log
cat gpurun_out/r2f_ops_gemm.log
LS_GEMM_SLAB_SINGLE=0 timeout 600 python tools/gemm_shapes.py --only conv --bns 160 > gpurun_out/r2f_shapes_ss0.txt 2>&1
LS_GEMM_SLAB_SINGLE=2 timeout 600 python tools/gemm_shapes.py --only conv --bns 160 > gpurun_out/r2f_shapes_ss2.txt 2>&1
paste <(grep -E "^conv|ctas=0" gpurun_out/r2f_shapes_ss0.txt) <(grep -E "^conv|ctas=0" gpurun_out/r2f_shapes_ss2.txt) | cut -c1-200
timeout 600 python bench.py --steps 3 --no-extras > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("fps", round(d["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3, r["other_kinds_ms_in_graph"])
PY
