#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_model.py -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -12
timeout 300 python tools/time_boundary.py 2>&1 | tail -2
