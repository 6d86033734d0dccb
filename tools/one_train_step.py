"""N eager training steps of the bench training leg, the last one inside the profiler range (for ncu launch lists).
usage: python tools/one_train_step.py [config] [batch] [warmup]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cenet_b200.train as T
from cenet_b200.networks import CENet
from oracle import fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "acdc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kw = fixtures.CONFIGS[name]
torch.manual_seed(1234)
m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.cuda().train()
eng = T.TrainEngine(m, "cuda:0", "bf16")
eng.use_graph = False
x = fixtures.synth_input(name, B).cuda()
y = torch.randint(0, kw["num_classes"], (B, 224, 224), device="cuda")
for _ in range(n):
    eng.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", eng.launches_per_step)
