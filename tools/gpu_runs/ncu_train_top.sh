#!/bin/bash
# full ncu captures of the largest training kernels inside the real step (ACDC, batch 24)
mkdir -p gpurun_out
CENET_B200_WGRAD_STREAM=0 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:"flash_bwd_dkv_kernel<\(int\)16|flash_bwd_dq_kernel<\(int\)16|flash_fwd_kernel<\(int\)16|flash_bwd_dkv_kernel<\(int\)64, \(int\)64, \(int\)0" -c 4 \
    -o gpurun_out/prof_train_flash python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_tf.log 2>&1
grep -c "flash" gpurun_out/ncu_tf.log; tail -n 1 gpurun_out/ncu_tf.log
timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "wgrad or gemm" 2>&1 | tail -3
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1 | cut -c1-160
