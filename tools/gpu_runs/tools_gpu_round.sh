#!/bin/bash
# One gpurun call: the driver's own test command, the training op breakdown, the bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -5 gpurun_out/gpu_tests_all.log
timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops.txt 2>&1
head -60 gpurun_out/train_ops.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -3 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"])
print(d["op_family_ms"]); print(d["roofline"]); 
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"], t.get("cpu_baseline"))
PY
