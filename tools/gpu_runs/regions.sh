#!/bin/bash
mkdir -p gpurun_out
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops_final.txt 2>&1
sed -n '/by region/,/by (op, region)/p' gpurun_out/train_ops_final.txt | head -40
