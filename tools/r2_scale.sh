#!/bin/bash
# strong / weak scaling lines at N GPUs: bash tools/r2_scale.sh N  (N = 1 runs without torchrun)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N"; fi
timeout 900 $TR bench.py --gpus $N --steps 10 --no-extras > gpurun_out/r2j_weak_${N}gpu.json 2> gpurun_out/r2j_weak_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --clip-segments 8 --steps 2 --no-extras > gpurun_out/r2j_clip8_${N}gpu.json 2> gpurun_out/r2j_clip8_${N}gpu.err
timeout 1500 $TR bench.py --gpus $N --clip-segments 94 --ddim-steps 50 --guidance 2.0 --no-extras > gpurun_out/r2j_clip94_${N}gpu.json 2> gpurun_out/r2j_clip94_${N}gpu.err
timeout 1500 $TR bench.py --gpus $N --clip-segments 94 --ddim-steps 50 --guidance 2.0 --segments-per-batch 2 --no-extras > gpurun_out/r2j_clip94_spb2_${N}gpu.json 2> gpurun_out/r2j_clip94_spb2_${N}gpu.err
python - <<PY
import json
for f in ("weak", "clip8", "clip94", "clip94_spb2"):
    try:
        d = json.loads(open(f"gpurun_out/r2j_{f}_${N}gpu.json").read().strip().splitlines()[-1])
        print(f, "N=$N", d["scaling"], "fps", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 1), d["clocks"]["sm_mhz"] if d["clocks"] else None)
    except Exception as e:
        print(f, "unreadable", e)
PY
