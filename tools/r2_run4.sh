#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_lin_one.py > gpurun_out/lin_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 2 -o gpurun_out/r2_lin_L0 -f python tools/gemm_lin_one.py > gpurun_out/lin_ncu.log 2>&1
tail -3 gpurun_out/lin_ncu.log
