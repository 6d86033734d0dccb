#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "dwconv" 2>&1 | tail -8
for s in 1; do
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 24 56 56 512 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 64 56 56 512 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 64 28 28 1024 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 24 14 14 1280 2>&1 | tail -1
  CENET_B200_DW_STAGED=$s python tools/one_dwconv.py 64 14 14 1280 2>&1 | tail -1
done
CENET_B200_DW_STAGED=0 python tools/one_dwconv.py 64 28 28 1024 2>&1 | tail -1
CENET_B200_DW_STAGED=0 python tools/one_dwconv.py 64 14 14 1280 2>&1 | tail -1
CENET_B200_DW_STAGED=1 ncu --set full --clock-control none --import-source on -k regex:"dwconv3x3_staged_kernel" -s 17 -c 1 \
   -o gpurun_out/prof_dw_staged2 python tools/one_dwconv.py 24 56 56 512 > gpurun_out/ncu_dw2.log 2>&1
tail -n 2 gpurun_out/ncu_dw2.log
