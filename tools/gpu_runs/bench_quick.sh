#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-skin512 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "launches_per_step", "eager_step_ms")})
r = d["roofline"]
print({k: r.get(k) for k in ("achieved", "frac", "ms_per_launch", "share_of_step")}, r.get("per_launch_view"))
for t in r.get("top_launches", []): print(t)
for t in r.get("top_excess", []): print("excess", t)
print(d["op_family_ms"])
PY
