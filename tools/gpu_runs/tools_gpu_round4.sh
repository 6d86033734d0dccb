#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hires.py -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -40
CENET_B200_WGRAD_STREAM=1 timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
