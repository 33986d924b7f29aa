#!/bin/bash
cd /root/repo
python -m pytest tests/test_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "layernorm" 2>&1 | tail -4
python -m pytest tests/test_model_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -6
LS_FOLD_LN=0 python tools/plan_gemm_times.py > gpurun_out/r2p_plan_gemm_fold0.txt 2>gpurun_out/r2p_pg0.err
LS_FOLD_LN=1 python tools/plan_gemm_times.py > gpurun_out/r2p_plan_gemm_fold1.txt 2>gpurun_out/r2p_pg1.err
tail -1 gpurun_out/r2p_plan_gemm_fold0.txt gpurun_out/r2p_plan_gemm_fold1.txt
