#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_model.py -q -m gpu -p no:cacheprovider --tb=short -k "dwconv or gather or model or graph or autograd or fp32 or bf16" 2>&1 | tail -15
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops3.txt 2>&1
grep "dwconv3x3_wgrad\|graph replay\|eager step" gpurun_out/train_ops3.txt
