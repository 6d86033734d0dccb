"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) into profiles/<tag>_launches.md and prints
selected counters of a full capture: python tools/summarise_ncu.py <tag>"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
leg = sys.argv[2] if len(sys.argv) > 2 else "infer"          # infer: launches.csv / train: launches_train.csv
src, what, script = {"infer": ("launches.csv", "one eager forward (Synapse config, batch 64, bf16)", "tools/profile_ncu.sh"),
                     "train": ("launches_train.csv", "one eager TRAINING step (ACDC config, batch 24, bf16: forward, Dice+CE, "
                               "backward, AdamW; includes the torch copy kernels of the per-step weight re-pack)",
                               "tools/profile_ncu_train.sh")}[leg]
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", src))) if len(r) > 10]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    v = float(r[ci["Metric Value"]].replace(",", "")); unit = r[ci["Metric Unit"]]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    short = re.sub(r"<unnamed>::|^void |\(.*", "", r[ci["Kernel Name"]])
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
out = [f"# ncu launch list, {what} -- {tag}", "",
       f"`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` around it ({script}). "
       "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
       f"total {tot/1e3:.3f} ms over {len(rows)-1} launches", "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:100]}` | {n} | {t:.1f} | {100*t/tot:.1f}% |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{tag}_launches{'' if leg == 'infer' else '_' + leg}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
