"""CPU debugging aid: TrainEngine on torch emulations vs oracle autograd (taps, loss, gradients)."""
import sys, os
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, fake_ops, fake_train_ops
import cenet_b200.train as T
T.ops, T.tops = fake_ops, fake_train_ops
from cenet_b200.networks import CENet
from oracle import cenet_oracle as O, fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "acdc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
size = int(sys.argv[3]) if len(sys.argv) > 3 else 64
flash = len(sys.argv) > 4 and sys.argv[4] == "flash"
kw = fixtures.CONFIGS[name]
torch.manual_seed(1234)
m = CENet(**kw); sd = fixtures.perturb_state(m.state_dict(), 1234); m.load_state_dict(sd); m.train()
eng = T.TrainEngine(m, "cpu", "fp32"); eng.use_graph = False; eng.use_flash = flash; eng.drop_path = False
x = fixtures.synth_input(name, B, size=size)
labels = torch.randint(0, kw["num_classes"], (B, size, size), generator=torch.Generator().manual_seed(5))
names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
taps_ref = {}
logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True, taps=taps_ref)
loss = O.criterion_dice_ce(logits, labels, kw["num_classes"])
grads = dict(zip(names, torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)))
eng.taps = {}
out = eng.train_step(x, labels, optimize=False)
for k, v in eng.taps.items():
    if k in taps_ref:
        r = taps_ref[k].detach()
        print(f"tap {k:32s} rel err {((v - r).norm() / r.norm()).item():.3e}")
lg = eng.buf("logits", logits.shape, torch.float32)
print("logits rel err", ((lg - logits.detach()).norm() / logits.detach().norm()).item(), "loss", out[0].item(), loss.item())
bad = 0
for k in names:
    g = grads[k]
    if g is None:
        print("unused", k, eng.GP[k].abs().max().item()); continue
    e = ((eng.GP[k] - g).norm() / max(g.norm().item(), 1e-12)).item()
    if e > 1e-3:
        bad += 1
        print(f"GRAD {k:60s} err {e:.3e} |g| {g.norm().item():.3e} |mine| {eng.GP[k].norm().item():.3e}")
print("bad grads:", bad, "of", len(names))
