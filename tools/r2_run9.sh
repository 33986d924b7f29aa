#!/bin/bash
mkdir -p gpurun_out
python tools/attn_temporal_one.py > gpurun_out/attn_one_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_one -s 3 -c 1 -o gpurun_out/r2_attn_one_temporal -f python tools/attn_temporal_one.py > gpurun_out/attn_one_ncu.log 2>&1
cat gpurun_out/attn_one_plain.log; tail -2 gpurun_out/attn_one_ncu.log
LS_ATTN_ONE=0 python tools/attn_temporal_one.py
HW=256 D=80 python tools/attn_temporal_one.py; LS_ATTN_ONE=0 HW=256 D=80 python tools/attn_temporal_one.py
