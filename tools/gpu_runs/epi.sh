#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/one_epilogue.py 2>&1 | tail -36
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py tests/test_gpu_model.py -x -q 2>&1 | tail -5
timeout 300 python tools/time_train.py 2>&1 | tail -1
bash tools/gpu_runs/bench_quick.sh 2>&1 | tail -12
