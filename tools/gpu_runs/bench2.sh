#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1200 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "e2e")})
print("train", {k: d["train"].get(k) for k in ("value", "ms_per_step", "e2e")})
print("synapse_dp", json.dumps(d.get("synapse_dp"), indent=1)[:1800])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
