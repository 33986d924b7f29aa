#!/bin/bash
# ncu --set full of the dominant kernel on its main-loop-bound shape (level-0 3x3 convolution, CTA pairs)
cd /root/repo
python tools/gemm_conv_one.py > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'gemm_tc' --launch-skip 3 -c 1 -f -o gpurun_out/r2u_gemm_conv python tools/gemm_conv_one.py > gpurun_out/r2u_gemm_conv_ncu.log 2>&1
tail -3 gpurun_out/r2u_gemm_conv_ncu.log
ncu -i gpurun_out/r2u_gemm_conv.ncu-rep --page raw --csv > gpurun_out/r2u_gemm_conv_raw.csv 2>/dev/null; wc -c gpurun_out/r2u_gemm_conv_raw.csv
