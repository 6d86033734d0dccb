#!/bin/bash
# Round-end evidence: the driver's own test command, the bench line, op breakdowns and ncu launch lists of both legs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -3 gpurun_out/gpu_tests_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -2 gpurun_out/bench_latest.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"])
print(d["op_family_ms"]); print(d["roofline"]); t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"]["value"])
r=json.loads([x for x in open("gpurun_out/bench_reference.json") if x.startswith("{")][0]); print("reference arm", r["value"], r["cpu_baseline"])
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops_final.txt 2>&1
head -14 gpurun_out/train_ops_final.txt
./tools/profile_ncu.sh > /dev/null 2>&1
CENET_B200_WGRAD_STREAM=0 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train.csv python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_launch_train.log 2>&1
tail -n 1 gpurun_out/ncu_launch_train.log
