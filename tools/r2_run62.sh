#!/bin/bash
cd /root/repo
timeout 500 python bench.py --clip-segments 94 --ddim-steps 50 --guidance 2.0 --steps 1 --warmup 1 --no-extras > gpurun_out/r2_final_clip94_1gpu.json 2> gpurun_out/r2_final_clip94_1gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_clip94_1gpu.json').read().strip().splitlines()[-1]); print('clip94 1gpu', d['value'], d['scaling'], d['ms_per_step'], d['e2e']['value'])"; tail -2 gpurun_out/r2_final_clip94_1gpu.err
