#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -6
for v in 0 1; do LS_LOOP_GRAPH=$v timeout 300 python bench.py --steps 4 --no-extras 2>gpurun_out/r2y_bench_lg$v.err | tee gpurun_out/r2y_bench_lg$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LOOP_GRAPH=$v fps', round(d['value'],2), 'ms/seg', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'unet_ms', round(d['unet_step_ms'],3), 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/r2y_bench_lg1.err
