#!/bin/bash
# per-launch device times of one training step (ACDC B=24): ncu launch list -> profiles/r2_launches_train.md
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_launch_train.log 2>&1; echo "ncu train rc=$?"
tail -3 gpurun_out/ncu_launch_train.log
python tools/summarise_ncu.py r2 train 2>&1 | tail -60
