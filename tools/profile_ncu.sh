#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one eager forward with its device time, (2) full capture of the
# dominant kernel.  Run under gpurun; summaries are produced afterwards by tools/summarise_ncu.py in the build container.
mkdir -p gpurun_out
CENET_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/one_forward.py synapse 64 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:diffattn_flash_kernel -s 1 -c 1 \
    -o gpurun_out/prof_diffattn python tools/one_diffattn.py 64 3136 128 8 > gpurun_out/ncu_da.log 2>&1
tail -2 gpurun_out/ncu_da.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 \
    -o gpurun_out/prof_gemm_fc1 python tools/one_gemm.py 200704 512 64 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
