"""Graph-replay time of the training step: python tools/time_train.py [config] [batch] [steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cenet_b200.train as T
from cenet_b200.networks import CENet
from oracle import fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "acdc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kw = fixtures.CONFIGS[name]
torch.manual_seed(1234)
m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.cuda().train()
eng = T.TrainEngine(m, "cuda:0", "bf16")
x = fixtures.synth_input(name, B).cuda()
y = torch.randint(0, kw["num_classes"], (B, 224, 224), device="cuda")
for _ in range(4):
    loss = eng.train_step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = eng.train_step(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
side = eng.side
print(f"{name} B={B} wgrad_stream={'on' if side is not None else 'off'}: {ms:.2f} ms/step -> {B / ms * 1e3:.1f} img/s; "
      f"launches/step {eng.launches_per_step}; loss {loss[0].item():.5f}"
      + (f"; side launches {side.launches} guard waits {side.waits}" if side is not None else ""))
