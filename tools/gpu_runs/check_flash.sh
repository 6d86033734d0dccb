#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "flash" 2>&1 | tail -3
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1 | cut -c1-160
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops7.txt 2>&1
grep "flash_bwd\|flash_fwd\|graph replay" gpurun_out/train_ops7.txt | head -14
