#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_whisper_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "geglu or gemm or whisper or encoder" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -s -k "stage2 or tiny_forward" 2>&1 | grep -i "rel-L2\|psnr\|passed\|failed\|error" | tail -12
timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2t_bench_gelu.err | tee gpurun_out/r2t_bench_gelu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('1-MUFU gelu fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3))"
python tools/plan_gemm_times.py 2>/dev/null | head -12
