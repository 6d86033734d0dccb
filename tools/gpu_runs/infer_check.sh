#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_org.py tests/test_gpu_hires.py tests/test_gpu_volume.py -x -q 2>&1 | tail -5
bash tools/gpu_runs/bench_quick.sh 2>&1 | grep -v "^{'op'\|^excess" | tail -6
