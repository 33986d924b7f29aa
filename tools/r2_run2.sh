#!/bin/bash
# round 2, GPU call 2: full GPU test suite on the new tree, bench (default + clip mode), VAE shape sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_gputests.log
cat gpurun_out/r2_gputests.log
timeout 900 python bench.py --steps 5 --profile-kernels > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 300 gpurun_out/r2b_bench.err
timeout 600 python bench.py --clip-segments 8 --no-extras > gpurun_out/r2b_bench_clip8.json 2> gpurun_out/r2b_bench_clip8.err
tail -c 300 gpurun_out/r2b_bench_clip8.err
timeout 600 python tools/gemm_shapes.py --only vae --bns 128,160,192,256 > gpurun_out/r2b_shapes_vae.txt 2>&1
python - <<'PY'
import json
for f in ("r2b_bench.json", "r2b_bench_clip8.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "unet_ms", round(d["unet_step_ms"], 3),
              "gemm frac", round(r.get("frac", 0), 3), "burst", (r.get("isolated_burst") or {}).get("frac"),
              "instep", (r.get("in_step_estimate") or {}).get("frac"), r.get("other_kinds_ms_in_graph"),
              "eager", d.get("gpu_eager_baseline"), "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
