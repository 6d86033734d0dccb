#!/bin/bash
./tools_gpu_tests.sh
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -2 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "eager_ms", round(d["eager_step_ms"],2), "launches", d["launches_per_step"], "clocks", d["clocks"])
print(d["op_family_ms"]); print(d["op_breakdown_ms"]); print(d["roofline"])
PY
