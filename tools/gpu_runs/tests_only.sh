#!/bin/bash
# gpurun -- bash tools/gpu_runs/tests_only.sh [pytest -k expression]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
