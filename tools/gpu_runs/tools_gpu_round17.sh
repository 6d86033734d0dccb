#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_model.py -q -m gpu -p no:cacheprovider --tb=short -k "rmsnorm or fea or model or fp32 or bf16 or graph or autograd or boundary" 2>&1 | tail -6
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1 | cut -c1-200
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops6.txt 2>&1
grep "fea_bwd\|diff_rmsnorm\|graph replay" gpurun_out/train_ops6.txt | head
