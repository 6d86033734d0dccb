#!/bin/bash
# ncu evidence for the training leg: (1) every launch of one eager training step with its device time,
# (2) full captures of the dominant training kernels (inside the real step, so shapes are the bench's).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train.csv python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_launch_train.log 2>&1
tail -2 gpurun_out/ncu_launch_train.log
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"flash_bwd_dkv_kernel|flash_bwd_dq_kernel|flash_fwd_kernel" -c 6 \
    -o gpurun_out/prof_train_flash python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_train_flash.log 2>&1
tail -2 gpurun_out/ncu_train_flash.log
