#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "dwconv" 2>&1 | tail -4
python tools/one_dwconv.py 24 56 56 512 2>&1 | tail -1
python tools/one_dwconv.py 64 56 56 512 2>&1 | tail -1
python tools/micro/hbm_rw.py 2>&1 | tail -3
timeout 300 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm.txt | tail -25
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -3 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_per_step"], "clocks", d["clocks"])
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"]["value"])
PY
