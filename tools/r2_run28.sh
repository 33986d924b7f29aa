#!/bin/bash
# LayerNorm fold: op tests, model tests, bench A/B
cd /root/repo
python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "layernorm or gemm or conv or geglu" 2>&1 | tail -8
python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -8
for v in 0 1; do LS_FOLD_LN=$v python bench.py --steps 3 --no-extras 2>gpurun_out/r2p_bench_fold$v.err | tee gpurun_out/r2p_bench_fold$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('FOLD=$v fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3), 'launches', r['launches_per_unet_forward'])"; done
