#!/bin/bash
mkdir -p gpurun_out
for c in relu; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_gemm_epi_${c}_v2 python tools/one_epi_case.py $c > gpurun_out/ncu_epi_$c.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_epi_$c.log
done
