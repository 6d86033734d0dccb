#!/bin/bash
for v in 0 1 0 1; do CENET_B200_WGRAD_STREAM=$v timeout 300 python tools/time_train.py acdc 24 30 2>&1 | tail -1 | cut -c1-110; done
