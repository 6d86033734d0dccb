#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py tests/test_gpu_train_model.py tests/test_gpu_model.py -x -q 2>&1 | tail -8
timeout 300 python tools/time_train.py 2>&1 | tail -2
