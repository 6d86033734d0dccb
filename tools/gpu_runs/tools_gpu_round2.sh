#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -5 gpurun_out/gpu_tests_all.log
./tools/profile_ncu_train.sh
