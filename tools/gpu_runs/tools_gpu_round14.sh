#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_model.py -q -m gpu -p no:cacheprovider --tb=short -k "loss or dice or boundary or diffattn" 2>&1 | tail -15
