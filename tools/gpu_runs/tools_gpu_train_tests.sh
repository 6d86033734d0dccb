#!/bin/bash
# GPU parity tests of the training kernels, one group per process (a sticky CUDA fault must not poison the others).
mkdir -p gpurun_out
: > gpurun_out/gpu_train_tests.log
for k in "gemm_training or a_mmajor or zout" "gemm_wgrad" "layernorm_bwd" "bn_stats or affine_act" "dwconv_wgrad or sumpool2" "flash" "softmax_bwd or lambda" "fea_bwd or layout" "ccu_train" "srm_train" "silu_mul" "resample or maxpool or head_upsample"; do
  echo "=== -k $k" >> gpurun_out/gpu_train_tests.log
  timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -m gpu --tb=short -p no:cacheprovider -k "$k" >> gpurun_out/gpu_train_tests.log 2>&1
done
for k in "fp32" "bf16" "graph_replay" "autograd_boundary"; do
  echo "=== model -k $k" >> gpurun_out/gpu_train_tests.log
  timeout 900 python -m pytest tests/test_gpu_train_model.py -q -m gpu --tb=short -p no:cacheprovider -k "$k" >> gpurun_out/gpu_train_tests.log 2>&1
done
grep -E "^===|passed|failed|^FAILED|^ERROR|Error|error:" gpurun_out/gpu_train_tests.log | cut -c1-250 | head -120
