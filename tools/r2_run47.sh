#!/bin/bash
cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; tail -c 400 gpurun_out/r2u_bench_2gpu.json; tail -3 gpurun_out/r2u_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --clip-segments 8 --steps 2 --warmup 1 > gpurun_out/r2u_clip8_2gpu.json 2> gpurun_out/r2u_clip8_2gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/r2u_clip8_2gpu.json').read().strip().splitlines()[-1]); print('clip8 2gpu', d['value'], d['scaling'], d['ms_per_step'])"
