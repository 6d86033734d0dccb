#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "diffattn or attn_tc or nonlocal or sr_attention" > gpurun_out/diffattn_tests.log 2>&1; echo "attention tests rc=$?"
tail -8 gpurun_out/diffattn_tests.log
timeout 600 python tools/one_diffattn_tc.py time 2>&1 | tail -6
timeout 300 python tools/one_attn_tc.py 2>&1 | tail -8
