#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "stem" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_hires.py -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -3
timeout 300 python tools/time_infer.py synapse 64 20 2>&1 | tail -1 | cut -c1-160
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1 | cut -c1-160
