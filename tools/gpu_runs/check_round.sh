#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -3 gpurun_out/gpu_tests_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -2 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_per_step"], "clocks", d["clocks"])
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
timeout 300 python tools/time_infer.py synapse 1 50 2>&1 | tail -1
timeout 300 python tools/time_infer.py acdc 1 50 2>&1 | tail -1
