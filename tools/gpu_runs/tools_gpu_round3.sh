#!/bin/bash
mkdir -p gpurun_out
CENET_B200_WGRAD_STREAM=0 timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
CENET_B200_WGRAD_STREAM=1 timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_train_model.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
