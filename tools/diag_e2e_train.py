"""Where does the end-to-end training step lose time at N>1?  torchrun --nproc-per-node N tools/diag_e2e_train.py"""
import contextlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import replicas
from cenet_b200.networks import CENet
from oracle import fixtures
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
kw = fixtures.CONFIGS["acdc"]
torch.manual_seed(1234)
with contextlib.redirect_stdout(sys.stderr):
    m = CENet(**kw)
m.load_state_dict(fixtures.perturb_state(m.state_dict(), 1234))
m = m.to(dev).train()
eng = m.train_engine(dev)
sync = replicas.GradSync(eng) if world > 1 else None
B = 24
x_host = fixtures.synth_input("acdc", B, 224, seed=100 + rank).pin_memory()
y_host = torch.randint(0, 4, (B, 224, 224), generator=torch.Generator().manual_seed(200 + rank)).pin_memory()
x_dev, y_dev = x_host.to(dev), y_host.to(dev)
xs, ys = torch.empty_like(x_dev), torch.empty_like(y_dev)
loss_host = torch.zeros(5).pin_memory()
for _ in range(4):
    eng.train_step(x_dev, y_dev)
def loop(name, body, n=10):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    replicas.barrier(dev)
    for e0, e1 in ev:
        e0.record(); body(); e1.record()
    replicas.barrier(dev)
    t = sum(a.elapsed_time(b) for a, b in ev) / n
    if rank == 0:
        print(f"{name}: {t:.2f} ms/step", flush=True)
def dev_only(): eng.train_step(x_dev, y_dev)
def e2e():
    xd = x_host.to(dev, non_blocking=True); yd = y_host.to(dev, non_blocking=True)
    loss = eng.train_step(xd, yd); loss_host.copy_(loss, non_blocking=True)
def e2e_prealloc():
    xs.copy_(x_host, non_blocking=True); ys.copy_(y_host, non_blocking=True)
    loss = eng.train_step(xs, ys); loss_host.copy_(loss, non_blocking=True)
def e2e_no_d2h():
    xs.copy_(x_host, non_blocking=True); ys.copy_(y_host, non_blocking=True)
    eng.train_step(xs, ys)
def h2d_only():
    xs.copy_(x_host, non_blocking=True); ys.copy_(y_host, non_blocking=True)
def dev_plus_d2h():
    loss = eng.train_step(x_dev, y_dev); loss_host.copy_(loss, non_blocking=True)
for name, fn in (("device only", dev_only), ("e2e (bench)", e2e), ("e2e prealloc", e2e_prealloc), ("e2e no d2h", e2e_no_d2h),
                 ("h2d only", h2d_only), ("device + d2h", dev_plus_d2h), ("device only again", dev_only)):
    loop(name, fn)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
