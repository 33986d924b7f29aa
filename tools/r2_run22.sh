#!/bin/bash
mkdir -p gpurun_out
LS_GEMM_EW16=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_boundary_gpu.py -q -m gpu -x --no-header -p no:cacheprovider -k "gemm or conv or pair" 2>&1 | tail -4
echo "EW16=0"; python tools/epi_ablate.py
echo "EW16=1"; LS_GEMM_EW16=1 python tools/epi_ablate.py
LS_GEMM_EW16=0 timeout 600 python tools/gemm_shapes.py --bns 160 > gpurun_out/r2m_shapes_ew8.txt 2>&1
LS_GEMM_EW16=1 timeout 600 python tools/gemm_shapes.py --bns 160 > gpurun_out/r2m_shapes_ew16.txt 2>&1
paste <(grep -E "ctas=0" gpurun_out/r2m_shapes_ew8.txt | awk '{print $3, $4}') <(grep -E "ctas=0" gpurun_out/r2m_shapes_ew16.txt | awk '{print $3, $4, $9}') <(grep -E "^[a-z]" gpurun_out/r2m_shapes_ew8.txt | cut -c1-60)
