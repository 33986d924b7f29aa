#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "upsample or conv3x3 or gemm_linear" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s 2>&1 | grep -i "rel-L2\|psnr\|passed\|failed\|error" | tail -30
for v in 0 1; do LS_UPCONV_FOLD=$v timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2w_bench_up$v.err | tee gpurun_out/r2w_bench_up$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('UPCONV_FOLD=$v fps', round(d['value'],2), 'ms/seg', round(d['ms_per_step'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3), 'launches', r['launches_per_unet_forward'])"; done
