#!/bin/bash
# gpurun -- bash tools/gpu_runs/check.sh : GPU parity suite + one default bench line (outputs under gpurun_out/)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "sustained", "latency_b1", "launches_per_step", "vs_eager", "cpu_baseline", "clocks")}
    keep["roofline_frac"] = (d.get("roofline") or {}).get("frac")
    keep["train"] = {k: (d.get("train") or {}).get(k) for k in ("value", "ms_per_step", "launches_per_step")}
    keep["eager_infer"] = (d.get("eager") or {}).get("infer"); keep["eager_train"] = (d.get("eager") or {}).get("train")
    keep["eager_err"] = (d.get("eager") or {}).get("error")
    keep["skin512"] = d.get("skin512")
    print(json.dumps(keep, indent=1)[:6000])
except Exception as e:
    print("parse failed", e)
PY
