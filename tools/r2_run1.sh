#!/bin/bash
# round 2, GPU call 1: A/B of the control-loop election (lane==0 vs elect.sync) + pair-mode sweep over real shapes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gemm or conv" -x --no-header -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r2_ops_gemm.log
cat gpurun_out/r2_ops_gemm.log
LS_SO_NAME=_C_ablate.so timeout 600 python tools/gemm_ablate.py > gpurun_out/r2_ablate_elect.txt 2>&1
LS_SO_NAME=_C_lane0_ablate.so QUICK=1 timeout 600 python tools/gemm_ablate.py > gpurun_out/r2_ablate_lane0.txt 2>&1
timeout 900 python tools/gemm_shapes.py > gpurun_out/r2_shapes_elect.txt 2>&1
LS_SO_NAME=_C_lane0.so timeout 600 python tools/gemm_shapes.py --bns 128,160,256 > gpurun_out/r2_shapes_lane0.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 3 --profile-kernels > gpurun_out/r2_bench_elect.json 2> gpurun_out/r2_bench_elect.err
LS_SO_NAME=_C_lane0.so timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2_bench_lane0.json 2> gpurun_out/r2_bench_lane0.err
tail -c 600 gpurun_out/r2_bench_elect.json; echo; tail -c 600 gpurun_out/r2_bench_lane0.json
