#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -x -q > gpurun_out/kernel_tests.log 2>&1; echo "kernel tests rc=$?"
tail -6 gpurun_out/kernel_tests.log
timeout 300 python tools/one_conv.py 2>&1 | tail -4
timeout 300 python tools/bench_gemm.py 2>&1 | tail -25
