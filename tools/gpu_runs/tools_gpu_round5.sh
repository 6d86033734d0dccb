#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_hires.py tests/test_gpu_train_model.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -40
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops2.txt 2>&1
head -24 gpurun_out/train_ops2.txt
