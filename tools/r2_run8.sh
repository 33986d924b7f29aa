#!/bin/bash
# round 2, GPU call 8: persistent one-tile tcgen05 attention (cross / short / packed temporal)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "attention" 2>&1 | tail -30 > gpurun_out/r2g_ops_attn.log
cat gpurun_out/r2g_ops_attn.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2g_model.log
cat gpurun_out/r2g_model.log
timeout 900 python bench.py --steps 5 --no-extras --profile-kernels > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), r["other_kinds_ms_in_graph"], "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3)
PY
grep "^unet attention" gpurun_out/r2g_bench.err
