"""Reference arms of bench.py: the UNMODIFIED reference (`baseline/_ref/src`, byte copy of /root/reference/src made by
tools/vendor_reference.py) run through its own public API -- `networks.CENet(**script kwargs)`, `utils.core.Criterion`,
`torch.optim.AdamW` -- on (a) the box's host cores, (b) the same B200 in PyTorch eager (cuDNN / cuBLAS / ATen), which is how
the reference's users actually run it and therefore the bar for the hand-written kernels.

Nothing of cenet_b200 is on these paths: no kernels, no engine, no module.  The only non-reference code is the timm / monai
stand-in (`oracle/ref_shim.py`, reached through tests/stubs) because neither package exists in this image, and
`oracle.fixtures` for the shared synthetic inputs / weight perturbation recipe.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")
SIZE = 224


def available():
    return os.path.isfile(os.path.join(REF_SRC, "networks", "__init__.py"))


def import_reference():
    """-> (networks module of the reference, utils.core module of the reference)"""
    if not available():
        raise RuntimeError("baseline/_ref not present: run `python tools/vendor_reference.py` where /root/reference exists")
    stubs = os.path.join(ROOT, "tests", "stubs")
    for p in (stubs, REF_SRC):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
        del sys.modules[k]
    with contextlib.redirect_stdout(sys.stderr):
        import networks                                            # the reference's own package (REF_SRC is first)
    assert os.path.realpath(networks.__file__).startswith(os.path.realpath(REF_SRC)), networks.__file__
    # utils/__init__.py drags thop / fvcore / matplotlib in; core.py itself only needs `flatten` from utils.py (unused by
    # dice / ce / boundary) -> load the unmodified file under a private parent package
    pkg = types.ModuleType("refutils")
    pkg.__path__ = []
    stub = types.ModuleType("refutils.utils")
    stub.flatten = None
    sys.modules["refutils"], sys.modules["refutils.utils"] = pkg, stub
    spec = importlib.util.spec_from_file_location("refutils.core", os.path.join(REF_SRC, "utils", "core.py"))
    core = importlib.util.module_from_spec(spec)
    sys.modules["refutils.core"] = core
    spec.loader.exec_module(core)
    return networks, core


def build_reference(nets, name, seed=1234):
    import torch
    from oracle import fixtures
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(sys.stderr):
        net = nets.CENet(**kw)
    net.load_state_dict(fixtures.perturb_state(net.state_dict(), seed), strict=True)
    return net, kw


# --------------------------------------------------------------------------------------------------------- host cores
def cpu_inference(name, batch, steps, warmup):
    """reference forward + `argmax(softmax(.,1),1)` (metrics_eval.py:49-52) on the host cores -> (slices/s, ms/step, cores)"""
    import torch
    from oracle import fixtures
    nets, _ = import_reference()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net, kw = build_reference(nets, name)
    net.eval()
    x = fixtures.synth_input(name, batch, SIZE)
    with torch.no_grad():
        for _ in range(warmup):
            torch.argmax(torch.softmax(net(x), dim=1), dim=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            torch.argmax(torch.softmax(net(x), dim=1), dim=1)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores


def cpu_train(name, batch, steps, warmup=0):
    """reference training iteration (main_acdc.py:241-262 without AMP): forward, Criterion('dice,ce'), backward, AdamW"""
    import torch
    from oracle import fixtures
    nets, core = import_reference()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net, kw = build_reference(nets, name)
    net.train()
    crit = core.Criterion(kw["num_classes"], types.SimpleNamespace(loss_type="dice,ce", loss_weights="0.5,0.5"))
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    x = fixtures.synth_input(name, batch, SIZE)
    y = torch.randint(0, kw["num_classes"], (batch, SIZE, SIZE), generator=torch.Generator().manual_seed(5)).float()

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores


# --------------------------------------------------------------------------------------------------------- same B200, eager
def _modes():
    import torch
    return {
        # stock: what main_*.py sets with its defaults (--deterministic 1: cudnn.deterministic, no benchmark; main_acdc.py:63-68)
        "fp32": dict(autocast=None, benchmark=False, deterministic=True, tf32=False),
        "bf16_autocast": dict(autocast=torch.bfloat16, benchmark=False, deterministic=True, tf32=False),
        # the strongest eager configuration we could find: cuDNN autotuning, TF32 matmuls, bf16 autocast
        "bf16_autocast_tuned": dict(autocast=torch.bfloat16, benchmark=True, deterministic=False, tf32=True),
    }


def _apply_mode(md):
    import torch
    torch.backends.cudnn.benchmark = md["benchmark"]
    torch.backends.cudnn.deterministic = md["deterministic"]
    torch.backends.cuda.matmul.allow_tf32 = md["tf32"]
    torch.backends.cudnn.allow_tf32 = True if md["tf32"] else torch.backends.cudnn.allow_tf32


def _ctx(md):
    import torch
    return torch.autocast("cuda", dtype=md["autocast"]) if md["autocast"] is not None else contextlib.nullcontext()


def gpu_eager(steps, warmup, infer_name="synapse", infer_batch=64, train_name="acdc", train_batch=24, dev="cuda:0",
              modes=None, fp16_amp_train=True):
    """Times the reference module in eager mode on `dev`.  Inference: forward + argmax(softmax) per batch (largest batch of
    64/32/16/8 that fits: the reference materialises 2h N x N fp32 maps per image, multihead_diffattn.py:92-124); `e2e` adds
    the H2D of the batch and the D2H of the label maps the way metrics_eval.py:48-53 does them (synchronous copies).
    Training: the loop body of main_acdc.py:241-262 with Criterion('dice,ce') and AdamW(lr 1e-4, wd 1e-4)."""
    import torch
    from oracle import fixtures
    nets, core = import_reference()
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
    out = {"device": torch.cuda.get_device_name(dev), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "infer": {}, "train": {}}
    all_modes = _modes()
    for mname in (modes or list(all_modes)):
        md = all_modes[mname]
        _apply_mode(md)
        # ---------------- inference ----------------
        net, kw = build_reference(nets, infer_name)
        net = net.to(dev).eval()
        res = None
        for B in [b for b in (64, 32, 16, 8) if b <= infer_batch]:
            try:
                x_host = fixtures.synth_input(infer_name, B, SIZE).pin_memory()
                x = x_host.to(dev)
                with torch.no_grad(), _ctx(md):
                    for _ in range(max(warmup, 3)):
                        lab = torch.argmax(torch.softmax(net(x), dim=1), dim=1)
                    torch.cuda.synchronize(dev)
                    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                    for e0, e1 in ev:
                        flush.zero_()
                        e0.record()
                        lab = torch.argmax(torch.softmax(net(x), dim=1), dim=1)
                        e1.record()
                    torch.cuda.synchronize(dev)
                    t_dev = sum(a.elapsed_time(b) for a, b in ev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        flush.zero_()
                        lab = torch.argmax(torch.softmax(net(x_host.to(dev, non_blocking=True)), dim=1), dim=1)
                        lab_host = lab.cpu()
                    e1.record()
                    torch.cuda.synchronize(dev)
                    t_e2e = e0.elapsed_time(e1)
                res = dict(batch=B, value=B * steps / (t_dev / 1e3), ms_per_step=t_dev / steps, e2e=B * steps / (t_e2e / 1e3),
                           unit="slices/s", peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
                del lab, lab_host, x
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                continue
        out["infer"][mname] = res
        del net
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    # ---------------- training ----------------
    tmodes = dict(all_modes)
    if fp16_amp_train:   # the scripts' real mode: `--amp` = fp16 autocast + GradScaler (main_acdc.py:192-199, 243-249)
        tmodes["fp16_amp_gradscaler"] = dict(autocast=torch.float16, benchmark=False, deterministic=True, tf32=False, scaler=True)
    for mname in (modes or list(tmodes)):
        md = tmodes[mname]
        _apply_mode(md)
        res = None
        for B in [b for b in (24, 12, 6) if b <= train_batch]:
            try:
                net, kw = build_reference(nets, train_name)
                net = net.to(dev).train()
                crit = core.Criterion(kw["num_classes"], types.SimpleNamespace(loss_type="dice,ce", loss_weights="0.5,0.5"))
                opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
                scaler = torch.amp.GradScaler() if md.get("scaler") else None
                x = fixtures.synth_input(train_name, B, SIZE, seed=100).to(dev)
                y = torch.randint(0, kw["num_classes"], (B, SIZE, SIZE), generator=torch.Generator().manual_seed(200)).float().to(dev)

                def step():
                    opt.zero_grad()
                    with _ctx(md):
                        loss = crit(net(x), y)
                    if scaler is not None:
                        scaler.scale(loss).backward()
                        scaler.step(opt)
                        scaler.update()
                    else:
                        loss.backward()
                        opt.step()
                    return loss
                for _ in range(max(warmup, 3)):
                    step()
                torch.cuda.synchronize(dev)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                for e0, e1 in ev:
                    flush.zero_()
                    e0.record()
                    loss = step()
                    e1.record()
                torch.cuda.synchronize(dev)
                t_dev = sum(a.elapsed_time(b) for a, b in ev)
                res = dict(batch=B, value=B * steps / (t_dev / 1e3), ms_per_step=t_dev / steps, unit="img/s",
                           final_loss=float(loss), peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
                del net, opt, x, y, loss
                break
            except torch.cuda.OutOfMemoryError:
                net = opt = None
                torch.cuda.empty_cache()
                continue
        out["train"][mname] = res
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    for leg in ("infer", "train"):
        ok = {k: v for k, v in out[leg].items() if v}
        if ok:
            best = max(ok, key=lambda k: ok[k]["value"])
            out[leg]["best"] = dict(mode=best, value=ok[best]["value"], batch=ok[best]["batch"],
                                    e2e=ok[best].get("e2e"))
    return out
