#!/bin/bash
mkdir -p gpurun_out
for spec in "wgrad_tc_kernel:60:3" "bn_bwd_partial_kernel:3:2" "bn_bwd_apply_kernel:3:2" "layernorm_bwd_kernel:40:2" "colsum_partial_kernel:100:2" "bn_stats_partial_kernel:3:2" "affine_act_kernel:3:2" "gemm_tc_kernel:250:3"; do
  k=${spec%%:*}; rest=${spec#*:}; sk=${rest%%:*}; c=${rest#*:}
  CENET_B200_WGRAD_STREAM=0 ncu --set full --clock-control none --profile-from-start off -k regex:"$k" -s $sk -c $c \
      -o gpurun_out/prof_tm_$k python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_tm_$k.log 2>&1
  tail -n 1 gpurun_out/ncu_tm_$k.log | cut -c1-100
done
