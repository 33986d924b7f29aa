#!/bin/bash
mkdir -p gpurun_out
LS_GEMM_H2=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -6
echo "H2=0"; python tools/epi_ablate.py
echo "H2=1"; LS_GEMM_H2=1 python tools/epi_ablate.py
LS_GEMM_H2=0 timeout 600 python tools/gemm_shapes.py --bns 160 --only "lin" > gpurun_out/r2n_shapes_h2_0.txt 2>&1
LS_GEMM_H2=1 timeout 600 python tools/gemm_shapes.py --bns 160 --only "lin" > gpurun_out/r2n_shapes_h2_1.txt 2>&1
for k in qkv ffout; do LS_GEMM_H2=0 python tools/gemm_shapes.py --bns 160 --only $k >> gpurun_out/r2n_shapes_h2_0.txt 2>&1; LS_GEMM_H2=1 python tools/gemm_shapes.py --bns 160 --only $k >> gpurun_out/r2n_shapes_h2_1.txt 2>&1; done
paste <(grep -E "ctas=0" gpurun_out/r2n_shapes_h2_0.txt | awk '{print $3, $4}') <(grep -E "ctas=0" gpurun_out/r2n_shapes_h2_1.txt | awk '{print $3, $4, $9}') <(grep -E "^[a-z]" gpurun_out/r2n_shapes_h2_0.txt | cut -c1-60)
LS_GEMM_H2=2 timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|PSNR|error" | tail -20
LS_GEMM_H2=2 timeout 600 python bench.py --steps 3 --no-extras > gpurun_out/r2n_bench_h2.json 2> gpurun_out/r2n_bench_h2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench_h2.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("H2=2: fps", round(d["value"], 2), "unet_ms", round(d["unet_step_ms"], 3), "gemm frac", round(r["frac"], 3), "gemm ms", r["avg_launch_us"] * r["launches_per_unet_forward"] / 1e3)
PY
