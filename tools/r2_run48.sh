#!/bin/bash
cd /root/repo
for v in 0 1; do LS_PDL=$v timeout 300 python bench.py --steps 3 --no-extras 2>gpurun_out/r2v_bench_pdl$v.err | tee gpurun_out/r2v_bench_pdl$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('PDL=$v fps', round(d['value'],2), 'unet_ms', round(d['unet_step_ms'],3), r['other_kinds_ms_in_graph'], 'gemm ms', round(r['avg_launch_us']*r['launches_per_unet_forward']/1e3,3), 'in-step gemm', round(r['in_step_estimate']['gemm_ms'],3))"; done
