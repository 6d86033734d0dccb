#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -q -m gpu -p no:cacheprovider --tb=short -k "dwconv" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train_model.py -x -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -8
timeout 300 python tools/time_train.py acdc 24 20 2>&1 | tail -1
CENET_B200_WGRAD_STREAM=0 timeout 600 python tools/profile_train_ops.py acdc 24 bf16 > gpurun_out/train_ops4.txt 2>&1
grep "dwconv3x3\|graph replay\|eager step" gpurun_out/train_ops4.txt | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/bench_infer.json") if x.startswith("{")][0])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
print(d["op_family_ms"])
PY
