#!/bin/bash
cd /root/repo
python tools/gn_parts_ablate.py 2>&1 | grep norm
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s 2>&1 | grep -i "rel-L2\|psnr\|passed\|failed\|error" | tail -40
timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2t_bench_tanh.err | tee gpurun_out/r2t_bench_tanh.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('tanh-silu fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3))"
timeout 600 python tools/c5_stress.py > gpurun_out/r2t_c5_stress_512px.json 2> gpurun_out/r2t_c5.err; tail -c 700 gpurun_out/r2t_c5_stress_512px.json
