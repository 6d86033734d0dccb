#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_all.log 2>&1
tail -3 gpurun_out/gpu_tests_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -3 gpurun_out/bench_latest.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_latest.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"])
print(d["op_family_ms"]); print(d["roofline"]); print(d["roofline_attention"])
t=d["train"]; print("train", t["value"], t["ms_per_step"], t["e2e"], t.get("cpu_baseline"))
PY
CENET_B200_GRAPH=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
   -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic.csv python tools/one_forward.py synapse 64 3 > gpurun_out/ncu_gemm_traffic.log 2>&1
tail -n 1 gpurun_out/ncu_gemm_traffic.log
