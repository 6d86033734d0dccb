#!/bin/bash
mkdir -p gpurun_out
for p in 0 1; do
  CENET_B200_PDL=$p timeout 300 python tools/time_infer.py synapse 64 20 2>&1 | tail -1
  CENET_B200_PDL=$p timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train_model.py -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -8
