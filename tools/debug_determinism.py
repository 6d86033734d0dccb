import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from cenet_b200.networks import CENet
name, B = sys.argv[1], int(sys.argv[2])
kw = fixtures.CONFIGS[name]
torch.manual_seed(1234)
m = CENet(**kw); m.load_state_dict(fixtures.perturb_state(m.state_dict())); m = m.cuda().eval()
x = fixtures.synth_input(name, B).cuda()
eng = m._engine(x)
def run():
    eng.taps = {}
    y = eng.forward(x).clone(); t = eng.taps; eng.taps = None
    # also grab internal buffers
    bufs = {k[1]: v.clone() for k, v in eng._bufs.items()}
    return y, t, bufs
y1, t1, b1 = run(); y2, t2, b2 = run(); y3, t3, b3 = run()
print("logits equal 1-2:", torch.equal(y1, y2), "2-3:", torch.equal(y2, y3))
for k in b1:
    e12 = not torch.equal(b1[k], b2[k]); e23 = not torch.equal(b2[k], b3[k])
    if e12 or e23:
        d = (b2[k].float() - b3[k].float()).abs()
        print(f"  DIFF {k:24s} shape {tuple(b1[k].shape)} 1-2:{e12} 2-3:{e23} maxabs(2-3) {d.max().item():.3e} n={int((d>0).sum())}")
