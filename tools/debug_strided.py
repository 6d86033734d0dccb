import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from cenet_b200.networks import CENet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kw = fixtures.CONFIGS["synapse"]
torch.manual_seed(1234)
m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.to("cuda").eval()
x = fixtures.synth_input("synapse", B).to("cuda")
eng = m._engine(x)
for graph in (False, True):
    eng.use_graph = graph
    ys = []
    for i in range(3):
        with torch.no_grad():
            ys.append(eng.forward(x).clone())
        torch.cuda.synchronize()
    print("graph", graph, "run-to-run max diff", [(ys[0] - y).abs().max().item() for y in ys[1:]])
eng.use_graph = False
eng.taps = {}
with torch.no_grad():
    eng.forward(x); t1 = {k: v.clone() for k, v in eng.taps.items()}
    eng.forward(x); t2 = {k: v.clone() for k, v in eng.taps.items()}
for k in t1:
    d = (t1[k].float() - t2[k].float()).abs().max().item()
    if d > 0: print("tap differs", k, d)
lab = m.predict(x)
y = m(x)
print("labels equal argmax(logits):", torch.equal(lab.cpu(), y.argmax(1).cpu()), (lab.cpu() != y.argmax(1).cpu()).sum().item())
