#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "groupnorm" 2>&1 | tail -3
python tools/gn_parts_ablate.py 2>&1 | grep producer
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "stage2 or bitwise or shared" 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2v_bench.err | tee gpurun_out/r2v_bench_short.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('f32x2 fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3))"
