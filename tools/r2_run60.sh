#!/bin/bash
cd /root/repo
python -m pytest tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2_final_boundary_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench_2gpu.json').read().strip().splitlines()[-1]); print('2gpu', d['value'], d['e2e']['value'], d['n_gpus'])"
