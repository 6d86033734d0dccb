#!/bin/bash
for p in 1 0 1; do
  CENET_B200_PDL=$p timeout 300 python tools/time_infer.py synapse 64 20 2>&1 | tail -1
  CENET_B200_PDL=$p timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
done
