#!/bin/bash
# new tcgen05 attention: direct tests under a short timeout first (a protocol bug must not hang the box), then timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "attn_tc or nonlocal_flash or sr_attention" > gpurun_out/attn_tc_tests.log 2>&1; echo "attn tests rc=$?"
tail -30 gpurun_out/attn_tc_tests.log
timeout 300 python tools/one_attn_tc.py 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -x -q -k "volume or model or hires" > gpurun_out/pytest_sel.log 2>&1; echo "selected tests rc=$?"
tail -15 gpurun_out/pytest_sel.log
