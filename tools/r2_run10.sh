#!/bin/bash
# round 2, GPU call 10 (2 GPUs): NCCL tests of the sharded clip entry + strong-scaling lines at N = 2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_boundary_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2h_boundary_2gpu.log
cat gpurun_out/r2h_boundary_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --clip-segments 8 --no-extras > gpurun_out/r2h_clip8_2gpu.json 2> gpurun_out/r2h_clip8_2gpu.err
tail -c 400 gpurun_out/r2h_clip8_2gpu.err
timeout 1200 $TR --master-port 29512 bench.py --gpus 2 --clip-segments 94 --ddim-steps 50 --guidance 2.0 --no-extras > gpurun_out/r2h_clip94_2gpu.json 2> gpurun_out/r2h_clip94_2gpu.err
tail -c 400 gpurun_out/r2h_clip94_2gpu.err
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --no-extras > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
python - <<'PY'
import json
for f in ("r2h_clip8_2gpu", "r2h_clip94_2gpu", "r2h_bench_2gpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["scaling"], "fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 1), d["config"]["workload"][:60])
    except Exception as e:
        print(f, "unreadable", e)
PY
