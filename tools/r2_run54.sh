#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s 2>&1 | grep -i "rel-L2\|psnr\|passed\|failed\|error" | tail -14
for v in 0 1; do LS_S2_INPLACE=$v timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2y_bench_s2$v.err | tee gpurun_out/r2y_bench_s2$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('S2_INPLACE=$v fps', round(d['value'],2), 'ms/seg', round(d['ms_per_step'],2), 'unet_ms', round(d['unet_step_ms'],3), 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3), 'launches', r['launches_per_unet_forward'])"; done
