"""Per-(op, region) device time of one eager TRAINING step (CUDA events around every launch group).
usage: python tools/profile_train_ops.py [config] [batch] [precision]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cenet_b200.train as T
from cenet_b200.engine import _TimedOps
from cenet_b200.networks import CENet
from oracle import fixtures
name = sys.argv[1] if len(sys.argv) > 1 else "acdc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
kw = fixtures.CONFIGS[name]
torch.manual_seed(1234)
m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.cuda().train()
eng = T.TrainEngine(m, "cuda:0", prec)
eng.use_graph = False
x = fixtures.synth_input(name, B).cuda()
y = torch.randint(0, kw["num_classes"], (B, 224, 224), device="cuda")
for _ in range(2):
    eng.train_step(x, y)
torch.cuda.synchronize()
real_ops, real_tops = T.ops, T.tops
class Tagged(_TimedOps):
    def __getattr__(self, name):
        if name in ("make_tables", "ACT_GELU_GRAD"):
            return getattr(self._inner, name)
        fn = super().__getattr__(name)
        if name != "gemm":
            return fn
        inner = getattr(self._inner, "gemm")

        def timed(a, w, out, **k):
            tc = (a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and not k.get("w_nmajor") and not k.get("a_mmajor")
                  and k.get("batch", 1) == 1 and k["ldw"] % 8 == 0 and (k.get("conv") is not None or (k["lda"] % 8 == 0 and k["K"] % 8 == 0
                  and (a.data_ptr() + 2 * k.get("a_off", 0)) % 16 == 0)) and k.get("impl", -1) != 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = inner(a, w, out, **k)
            e1.record()
            self.records.append(("gemm.tc" if tc else f"gemm.simt[b{k.get('batch', 1) > 1}]", self.tag, e0, e1))
            return r
        return timed
t1 = Tagged(real_ops)
class Tagged2(Tagged):
    @property
    def tag(self):
        return t1.tag
    @tag.setter
    def tag(self, v):
        pass
t2 = Tagged2(real_tops)
T.ops, T.tops = t1, t2
steps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.train_step(x, y)
e1.record()
torch.cuda.synchronize()
T.ops, T.tops = real_ops, real_tops
agg = {}
for prox in (t1, t2):
    for (op, tag), (ms, n) in prox.summary().items():
        a = agg.setdefault(op, [0.0, 0]); a[0] += ms / steps; a[1] += n // steps
fine = {}
for prox in (t1, t2):
    for (op, tag), (ms, n) in prox.summary().items():
        fine[(op, tag)] = (ms / steps, n // steps)
tot = sum(v[0] for v in agg.values())
print(f"config {name} B={B} {prec}: eager step (events) {e0.elapsed_time(e1)/steps:.2f} ms; sum of op times {tot:.2f} ms; "
      f"{sum(v[1] for v in agg.values())} op calls")
for op, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{n:<4d} {op}")
print("---- by region (eager op time, launches)")
reg = {}
for (op, tag), (ms, n) in fine.items():
    a = reg.setdefault(tag, [0.0, 0]); a[0] += ms; a[1] += n
for tag, (ms, n) in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{n:<4d} {tag}")
print("---- by (op, region), top 45")
for (op, tag), (ms, n) in sorted(fine.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{n:<4d} {op}@{tag}")
# graph replay timing
eng.use_graph = True
for _ in range(3):
    eng.train_step(x, y)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    eng.train_step(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"graph replay: {ms:.2f} ms/step -> {B / ms * 1e3:.1f} img/s; launches/step {eng.launches_per_step}; "
      f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
