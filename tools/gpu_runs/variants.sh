#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_org.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_train_model.py -x -q -k "fp32_matches" 2>&1 | tail -5
