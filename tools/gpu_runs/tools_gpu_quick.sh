#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=short -p no:cacheprovider -k "gemm or conv_same" 2>&1 | tail -5
timeout 300 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm.txt
