#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -2 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"])
print(d["roofline"]["frac"], d["roofline_attention"]["frac"], d["roofline_attention"]["mufu_bound"]["frac"])
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"]["value"], t["cpu_baseline"]); print(d["cpu_baseline"])
PY
