#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "wgrad" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_train_model.py -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -3
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1 | cut -c1-160
