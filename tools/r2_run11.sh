#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/gather_probe.py 2>&1 | grep "rank" | grep pass | tee gpurun_out/r2i_gather_probe.txt
nvidia-smi topo -m 2>&1 | head -8
