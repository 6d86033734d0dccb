#!/bin/bash
# Runs the GPU parity tests one group per process (a sticky CUDA fault in one group must not poison the others).
mkdir -p gpurun_out
: > gpurun_out/gpu_tests.log
for k in gemm_plain gemm_epilogue conv_same "conv_strided or batched" "layernorm or dwconv or layout or stem" sr_attention diffattn nonlocal "fea or ccu" "head or dice"; do
  echo "=== -k $k" >> gpurun_out/gpu_tests.log
  timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=short -p no:cacheprovider -k "$k" >> gpurun_out/gpu_tests.log 2>&1
done
for k in fp32_precision bf16_precision "golden or graph or other"; do
  echo "=== model -k $k" >> gpurun_out/gpu_tests.log
  timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -p no:cacheprovider -k "$k" >> gpurun_out/gpu_tests.log 2>&1
done
grep -E "^===|passed|failed|^FAILED|^ERROR" gpurun_out/gpu_tests.log
