#!/bin/bash
# gpurun -- bash tools/gpu_runs/r2_iter.sh : new-kernel parity + micro timings + quick bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -m gpu -x -q -k "mixffn or diffattn_fwd_train or flash_fwd" > gpurun_out/pytest_sel.log 2>&1; echo "pytest sel rc=$?"
tail -15 gpurun_out/pytest_sel.log
for a in "64 56 56 512 64" "64 28 28 1024 128" "16 128 128 512 64" "16 64 64 1024 128"; do timeout 120 python tools/one_mixffn.py $a; done 2>&1 | tee gpurun_out/one_mixffn.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train_model.py -m gpu -x -q > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?"
tail -8 gpurun_out/pytest_model.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-skin512 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "launches_per_step", "eager_step_ms")})
r = d["roofline"]
print({k: r.get(k) for k in ("achieved", "frac", "ms_per_launch", "share_of_step")}, r.get("per_launch_view"))
print(d["op_family_ms"])
print({k: (d.get("train") or {}).get(k) for k in ("value", "ms_per_step", "launches_per_step")})
PY
