"""Time of the UNCHANGED reference loop on the drop-in module (net(x); criterion; backward; torch AdamW) vs the fused step."""
import sys, os, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200.networks import CENet
from cenet_b200.losses import Criterion
from oracle import fixtures
kw = fixtures.CONFIGS["acdc"]
torch.manual_seed(1234)
m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.cuda().train()
B = 24
x = fixtures.synth_input("acdc", B).cuda()
y = torch.randint(0, 4, (B, 224, 224), device="cuda")
crit = Criterion(4, types.SimpleNamespace(loss_type="dice,ce", loss_weights="0.5,0.5"))
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
def step():
    loss = crit(m(x), y.float())
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss
for _ in range(4):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"reference loop on the drop-in module (graphs {'on' if m.train_engine('cuda:0')._ab_graphs else 'off'}): {ms:.2f} ms/step -> "
      f"{B / ms * 1e3:.1f} img/s; loss {loss.item():.4f}")
