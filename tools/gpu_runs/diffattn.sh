#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "diffattn" > gpurun_out/diffattn_tests.log 2>&1; echo "diffattn tests rc=$?"
tail -5 gpurun_out/diffattn_tests.log
timeout 600 python tools/one_diffattn_tc.py time 2>&1 | tail -6
