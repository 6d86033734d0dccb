#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
CENET_B200_GRAPH=0 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tc_kernel" -c 44 \
    -o /tmp/ncu/prof_infer_gemm python tools/one_forward.py synapse 64 3 > gpurun_out/ncu_ig.log 2>&1
tail -n 1 gpurun_out/ncu_ig.log | cut -c1-100
ncu -i /tmp/ncu/prof_infer_gemm.ncu-rep --page raw --csv > gpurun_out/infer_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/infer_gemm_raw.csv
