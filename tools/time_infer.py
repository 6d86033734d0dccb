"""Graph-replay time of the inference forward: python tools/time_infer.py [config] [batch] [steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "synapse"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
bench.CONFIG_NAME = name
m, sd, kw = bench.build_model()
m = m.cuda().eval()
x = fixtures.synth_input(name, B).cuda()
eng = m._engine(x)
for _ in range(4):
    lab = eng.forward(x, labels=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    lab = eng.forward(x, labels=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{name} B={B}: {ms:.3f} ms/forward -> {B / ms * 1e3:.1f} slices/s; launches {eng.launches_per_forward}; "
      f"label checksum {int(lab.sum())}")
