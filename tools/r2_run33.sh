#!/bin/bash
cd /root/repo
python tools/gn_parts_ablate.py 2>&1 | grep norm
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -15
for v in 0 1; do LS_GN_PARTS=$v timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2q_bench_gnp$v.err | tee gpurun_out/r2q_bench_gnp$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('GN_PARTS=$v fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3), 'launches', r['launches_per_unet_forward'])"; done
