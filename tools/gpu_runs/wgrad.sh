#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/one_wgrad.py 10 2>&1 | tail -40 | tee gpurun_out/one_wgrad.txt
timeout 300 python tools/time_train.py 2>&1 | tail -5
