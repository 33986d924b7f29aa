#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "groupnorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s -k "stage2 or tiny or bitwise or shared or vae" 2>&1 | grep -i "rel-L2\|psnr\|passed\|failed\|error" | tail -16
for v in 0 1; do LS_GN_GROUP=$v timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2y_bench_gng$v.err | tee gpurun_out/r2y_bench_gng$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('GN_GROUP=$v fps', round(d['value'],2), 'ms/seg', round(d['ms_per_step'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'])"; done
