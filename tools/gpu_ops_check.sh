#!/bin/bash
# Runs each op-test group in its own process (a trapping kernel poisons the CUDA context of its process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/ops_smi.txt 2>&1
for k in gemm_linear gemm_bias gemm_geglu gemm_batched conv3x3 conv_stride2 upsample groupnorm layernorm softmax "attention and not temporal" attention_temporal concat13 layout; do
  echo "=== $k" >> gpurun_out/ops.log
  timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$k" --no-header -p no:cacheprovider 2>&1 | tail -40 >> gpurun_out/ops.log
done
grep -E "^===|passed|failed|error" gpurun_out/ops.log
