#!/bin/bash
# round 2, GPU call 16: time path + audio K/V hoisted out of the per-step graph: full GPU suite + headline bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r2k_gputests.log
cat gpurun_out/r2k_gputests.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-kernels > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), "burst", round(r["isolated_burst"]["frac"], 3), "instep", round(r["in_step_estimate"]["frac"], 3), "whole", round(r["whole_step"]["frac"], 3), r["other_kinds_ms_in_graph"], "traffic", r["traffic"], "launches", d["gpu_launches"], d["c_abi_calls_counted"])
print("eager", d["gpu_eager_baseline"]["value"], d["gpu_eager_baseline"]["unet_step_ms"], "cpu", d["cpu_baseline"]["value"], "pixels", d["from_pixels"]["value"])
PY
python __graft_entry__.py smoke 2>&1 | tail -2
