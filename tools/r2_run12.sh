#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29540
for v in A B; do
  if [ $v = B ]; then export LS_BENCH_NO_SAMPLER=1; fi
  P=$((P+1))
  timeout 900 $TR --master-port $P bench.py --gpus 2 --steps 10 --no-extras 2>gpurun_out/r2i_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'fps', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks'])"
  tail -c 300 gpurun_out/r2i_$v.err
done
