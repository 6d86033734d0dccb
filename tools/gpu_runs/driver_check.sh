#!/bin/bash
# what the driver runs at round end: the GPU test suite in one process, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -3 gpurun_out/gpu_tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
