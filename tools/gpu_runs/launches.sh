#!/bin/bash
# per-launch device times of one forward (B=64 Synapse): ncu launch list (cold-cache, serialised) -> profiles/r2_launches.md
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/one_forward.py synapse 64 3 > gpurun_out/launches_infer.log 2>&1; echo "ncu infer rc=$?"
python tools/summarise_ncu.py r2 infer 2>&1 | tail -45
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10]
hdr=rows[0]; ci={h:i for i,h in enumerate(hdr)}
# gemm_tc launches sorted by duration
g=[]
for r in rows[1:]:
    if "gemm_tc_kernel" in r[ci["Kernel Name"]]:
        v=float(r[ci["Metric Value"]].replace(",","")); u=r[ci["Metric Unit"]]
        v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
        g.append((v, r[ci["Grid Size"]], r[ci["Block Size"]]))
print("gemm_tc launches:", len(g), "total us", sum(x[0] for x in g))
import statistics
print("median us", statistics.median(x[0] for x in g), "min", min(x[0] for x in g), "max", max(x[0] for x in g))
hist=collections.Counter(int(x[0]//5)*5 for x in g)
print(sorted(hist.items()))
PY
