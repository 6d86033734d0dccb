#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 2 3 4; do
  echo "== CENET_DA_POLY=$p"
  CENET_DA_POLY=$p python tools/one_diffattn.py 64 3136 128 8 2>&1 | tail -1
  CENET_DA_POLY=$p timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=line -k "diffattn" 2>&1 | tail -2
done
