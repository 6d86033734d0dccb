#!/bin/bash
mkdir -p gpurun_out
for s in 0 1; do
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 24 56 56 512 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 64 56 56 512 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 24 14 14 1280 2>&1 | tail -1
done
CENET_B200_DW_STAGED=1 ncu --set full --clock-control none --import-source on -k regex:"dwconv3x3_staged_kernel|dw_wgrad_staged_kernel" -s 8 -c 3 \
   -o gpurun_out/prof_dw_staged python tools/one_dwconv.py 24 56 56 512 > gpurun_out/ncu_dw1.log 2>&1
CENET_B200_DW_STAGED=0 ncu --set full --clock-control none --import-source on -k regex:"dwconv3x3_kernel|dw_wgrad_partial_kernel" -s 8 -c 3 \
   -o gpurun_out/prof_dw_reg python tools/one_dwconv.py 24 56 56 512 > gpurun_out/ncu_dw0.log 2>&1
tail -2 gpurun_out/ncu_dw1.log gpurun_out/ncu_dw0.log
