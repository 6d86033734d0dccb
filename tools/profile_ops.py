"""Per-(op, region) device time of one eager forward (CUDA events around every launch)."""
import sys, os, json, torch, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "synapse"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bench.CONFIG_NAME = name
from oracle import fixtures
m, sd, kw = bench.build_model()
m = m.cuda().eval()
x = fixtures.synth_input(name, B).cuda()
eng = m._engine(x)
prof = eng.profile_ops(x, labels=True, steps=3)
tot = sum(v[0] for v in prof.values())
rows = sorted(prof.items(), key=lambda kv: -kv[1][0])
print(f"config {name} B={B}: eager step {tot:.3f} ms, {sum(v[1] for v in prof.values())} launches")
for (op, tag), (ms, n) in rows:
    print(f"{ms:8.4f} ms {100*ms/tot:5.1f}%  x{n:<3d} {op}@{tag}")
