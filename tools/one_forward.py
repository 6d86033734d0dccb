"""N eager forwards of the bench workload (for ncu launch lists)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import fixtures
name, B, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
bench.CONFIG_NAME = name
m, sd, kw = bench.build_model()
m = m.cuda().eval()
x = fixtures.synth_input(name, B).cuda()
eng = m._engine(x)
for _ in range(n):
    eng.forward(x, labels=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one forward is captured
eng.forward(x, labels=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per forward:", eng.launches_per_forward)
