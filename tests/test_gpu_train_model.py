"""Whole-model TRAINING parity on the B200: one step of the hand-written forward + backward (cenet_b200.train.TrainEngine,
all kernels through the C ABI) against autograd through the CPU oracle in train mode (batch-statistics BatchNorm, DropPath
off, Dice+CE with weights 0.5/0.5) on identical weights / inputs / labels.
  precision="fp32": loss and every parameter gradient to 1e-3 (logic);
  precision="bf16" (product path): loss to 1e-2, gradients by direction (cosine) and norm -- bf16 rounding of activations
  accumulates through ~100 layers, so per-tensor tolerances are looser and written below.
Also: CUDA-graph replay == eager, run-to-run determinism, the autograd drop-in boundary (`net(x)`; `loss.backward()`) and
that a few AdamW steps reduce the loss."""
import json
import os

import pytest
import torch

from conftest import ROOT
from oracle import cenet_oracle as O
from oracle import fixtures

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_CACHE = {}


def _ref(name, batch, size):
    key = (name, batch, size)
    if key not in _CACHE:
        from cenet_b200.networks import CENet
        kw = fixtures.CONFIGS[name]
        torch.manual_seed(1234)
        m = CENet(**kw)
        sd = fixtures.perturb_state(m.state_dict(), 1234)
        x = fixtures.synth_input(name, batch, size=size)
        labels = torch.randint(0, kw["num_classes"], (batch, size, size), generator=torch.Generator().manual_seed(5))
        names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
        leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
        logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True)
        loss = O.criterion_dice_ce(logits, labels, kw["num_classes"])
        grads = dict(zip(names, torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)))
        _CACHE[key] = (kw, sd, x, labels, loss.item(), logits.detach(), grads)
    return _CACHE[key]


def _engine(name, precision, sd):
    from cenet_b200.networks import CENet
    from cenet_b200.train import TrainEngine
    m = CENet(**fixtures.CONFIGS[name])
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    eng = TrainEngine(m, DEV, precision)
    eng.drop_path = False
    return m, eng


def _dump(tag, rows):
    path = os.path.join(ROOT, "gpurun_out", f"train_grad_errors_{tag}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(rows, open(path, "w"), indent=1)


@pytest.mark.parametrize("name,batch,size", [("acdc", 2, 224), ("synapse", 2, 96), ("skin", 2, 96), ("acdc_b1", 2, 96),
                                             ("acdc_add", 2, 96), ("synapse_uprb", 2, 64), ("acdc_uptc", 2, 64)])   # (uprb at 96: a max-pool near-tie makes even the fp32 emulation differ by 3e-3)
def test_train_step_fp32_matches_oracle_autograd(name, batch, size):
    kw, sd, x, labels, loss_ref, logits_ref, gref = _ref(name, batch, size)
    m, eng = _engine(name, "fp32", sd)
    eng.use_graph = False
    out = eng.train_step(x.to(DEV), labels.to(DEV), optimize=False)
    torch.cuda.synchronize()
    assert abs(out[0].item() - loss_ref) < 1e-4 * max(1.0, abs(loss_ref)), (out[0].item(), loss_ref)
    rows, bad = {}, []
    gn = max(g.norm().item() for g in gref.values() if g is not None)
    for k, g in gref.items():
        mine = eng.GP[k].cpu()
        if g is None:
            assert mine.abs().max().item() == 0.0
            continue
        err = (mine - g).norm().item()
        rows[k] = err / max(g.norm().item(), 1e-12)
        if not err < 2e-3 * g.norm().item() + 1e-6 * gn:
            bad.append((k, rows[k], g.norm().item()))
    _dump(f"fp32_{name}", rows)
    assert not bad, bad[:20]


@pytest.mark.parametrize("name,batch,size", [("acdc", 2, 224), ("synapse", 2, 224)])
def test_train_step_bf16_matches_oracle_autograd(name, batch, size):
    kw, sd, x, labels, loss_ref, logits_ref, gref = _ref(name, batch, size)
    m, eng = _engine(name, "bf16", sd)
    eng.use_graph = False
    out = eng.train_step(x.to(DEV), labels.to(DEV), optimize=False)
    torch.cuda.synchronize()
    lg = eng.buf("logits", logits_ref.shape, torch.float32).cpu()
    rel_logits = ((lg - logits_ref).norm() / logits_ref.norm()).item()
    assert rel_logits < 2e-2, rel_logits
    assert abs(out[0].item() - loss_ref) < 1e-2 * max(1.0, abs(loss_ref)), (out[0].item(), loss_ref)
    rows, cos_w, tot = {}, 0.0, 0.0
    flat_m, flat_r = [], []
    for k, g in gref.items():
        if g is None:
            continue
        mine = eng.GP[k].cpu().flatten()
        gg = g.flatten()
        rows[k] = ((mine - gg).norm() / max(gg.norm().item(), 1e-12)).item()
        flat_m.append(mine)
        flat_r.append(gg)
    _dump(f"bf16_{name}", rows)
    fm, fr = torch.cat(flat_m), torch.cat(flat_r)
    cos = torch.dot(fm, fr) / (fm.norm() * fr.norm())
    assert torch.isfinite(fm).all()
    assert cos.item() > 0.99, cos.item()                                  # direction of the full gradient
    assert abs(fm.norm().item() / fr.norm().item() - 1.0) < 0.05           # and its length
    big = [k for k, g in gref.items() if g is not None and g.norm().item() > 1e-3 * fr.norm().item()]
    worst = max(rows[k] for k in big)
    assert worst < 0.25, (worst, [k for k in big if rows[k] == worst])
    med = sorted(rows[k] for k in big)[len(big) // 2]
    assert med < 0.05, med


def test_graph_replay_determinism_and_loss_decreases():
    kw, sd, x, labels, *_ = _ref("acdc", 2, 224)
    m, eng = _engine("acdc", "bf16", sd)
    xd, ld = x.to(DEV), labels.to(DEV)
    losses = []
    for i in range(6):
        out = eng.train_step(xd, ld, lr=2e-4)
        losses.append(out[0].item())
    assert eng.launches_per_step and eng.launches_per_step > 500
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0], losses
    # same state, same data -> bit-identical step (no float atomics anywhere)
    m2, eng2 = _engine("acdc", "bf16", sd)
    m3, eng3 = _engine("acdc", "bf16", sd)
    eng3.use_graph = False
    a = [eng2.train_step(xd, ld, lr=2e-4)[0].item() for _ in range(3)]
    b = [eng3.train_step(xd, ld, lr=2e-4)[0].item() for _ in range(3)]
    assert a == b == losses[:3], (a, b, losses[:3])
    assert torch.equal(eng2.pflat, eng3.pflat)


def test_autograd_boundary_matches_fused_step():
    """`net(x)` in train() mode + torch's own loss / backward / AdamW (the reference loop) vs the fused engine step"""
    kw, sd, x, labels, loss_ref, *_ = _ref("acdc", 2, 224)
    from cenet_b200.networks import CENet
    m = CENet(**fixtures.CONFIGS["acdc"])
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    m.train_engine(DEV).drop_path = False
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
    logits = m(x.to(DEV))
    assert logits.requires_grad and logits.shape == (2, 4, 224, 224)
    loss = O.criterion_dice_ce(logits, labels.to(DEV), 4)
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert abs(loss.item() - loss_ref) < 1e-2 * max(1.0, abs(loss_ref))
    m2, eng2 = _engine("acdc", "bf16", sd)
    eng2.use_graph = False
    eng2.train_step(x.to(DEV), labels.to(DEV), lr=1e-4, weight_decay=1e-4)
    g1 = torch.cat([p.grad.flatten() for p in m.parameters()])
    g2 = torch.cat([eng2.GP[n].flatten() for n, _ in m2.named_parameters()])
    assert ((g1 - g2).norm() / g2.norm()).item() < 1e-3                    # torch's Dice+CE backward vs the fused loss kernel
    p1 = torch.cat([p.detach().flatten() for p in m.parameters()])
    p2 = torch.cat([eng2.P[n].flatten() for n, _ in m2.named_parameters()])
    assert (p1 - p2).abs().max().item() < 2.1e-4                           # one AdamW step moves each weight by <= lr
    # iterations 2..4 of the reference loop: the second call captures forward / backward as CUDA graphs, later ones replay
    for it in range(3):
        loss = O.criterion_dice_ce(m(x.to(DEV)), labels.to(DEV), 4)
        opt.zero_grad()
        loss.backward()
        opt.step()
        out2 = eng2.train_step(x.to(DEV), labels.to(DEV), lr=1e-4, weight_decay=1e-4)
        assert abs(loss.item() - out2[0].item()) < 5e-3 * max(1.0, abs(out2[0].item())), (it, loss.item(), out2[0].item())
    st = m.train_engine(DEV)._ab_graphs[(2, 224, 224)]
    assert "fwd" in st and "bwd" in st                                     # the graph path really ran
    p1 = torch.cat([p.detach().flatten() for p in m.parameters()])
    p2 = torch.cat([eng2.P[n].flatten() for n, _ in m2.named_parameters()])
    assert (p1 - p2).abs().max().item() < 8.1e-4
    # eval after training uses the updated weights and running statistics through the inference engine
    m.eval()
    with torch.no_grad():
        y = m(x.to(DEV))
    assert torch.isfinite(y).all()


def test_boundary_criterion_fused_step_and_dropin_module(monkeypatch):
    """`--loss_type boundary` (the ACDC / Synapse script default, acdc.sh:63): the fused step with the Boundary-DoU term
    against oracle autograd, and the drop-in `losses.Criterion` on the autograd boundary of the module."""
    import types
    from cenet_b200.losses import Criterion
    name, batch, size = "acdc", 2, 96
    kw = fixtures.CONFIGS[name]
    from cenet_b200.networks import CENet
    torch.manual_seed(1234)
    sd = fixtures.perturb_state(CENet(**kw).state_dict(), 1234)
    x = fixtures.synth_input(name, batch, size=size)
    coarse = torch.randint(0, 4, (batch, 1, size // 8, size // 8), generator=torch.Generator().manual_seed(5)).float()
    labels = torch.nn.functional.interpolate(coarse, scale_factor=8, mode="nearest")[:, 0].long()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    loss_ref = O.criterion(O.cenet_forward(leaf, O.Cfg(**kw), x, training=True), labels, 4, w_boundary=1.0)
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [leaf[k] for k in names], allow_unused=True)))
    monkeypatch.setenv("CENET_B200_PRECISION", "fp32")      # module.train_engine() then hands out the fp32 validation plan
    m = CENet(**kw)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    eng = m.train_engine(DEV)
    assert eng.precision == "fp32"
    eng.drop_path = False
    eng.use_graph = False
    out = eng.train_step(x.to(DEV), labels.to(DEV), w_dice=0.0, w_ce=0.0, w_boundary=1.0, optimize=False)
    torch.cuda.synchronize()
    assert abs(out[0].item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    gn = max(g.norm().item() for g in gref.values() if g is not None)
    bad = [(k, (eng.GP[k].cpu() - g).norm().item() / max(g.norm().item(), 1e-12)) for k, g in gref.items()
           if g is not None and not (eng.GP[k].cpu() - g).norm().item() < 2e-3 * g.norm().item() + 1e-6 * gn]
    assert not bad, bad[:10]
    # drop-in Criterion through autograd: same loss, same gradients as the fused step
    crit = Criterion(4, types.SimpleNamespace(loss_type="boundary", loss_weights="1.0"))
    fused = eng.gflat.clone()
    m.zero_grad(set_to_none=True)
    loss = crit(m(x.to(DEV)), labels.to(DEV).float())
    loss.backward()
    assert abs(loss.item() - out[0].item()) < 1e-5
    g_auto = torch.cat([p.grad.flatten() for _, p in sorted(m.named_parameters(), key=lambda np_: eng.param_offsets[np_[0]])])
    g_fused = torch.cat([fused[eng.param_offsets[n]:eng.param_offsets[n] + p.numel()]
                         for n, p in sorted(m.named_parameters(), key=lambda np_: eng.param_offsets[np_[0]])])
    assert ((g_auto - g_fused).norm() / g_fused.norm()).item() < 1e-5


def test_benchmarked_batch_24_bf16_step_matches_fp32_step():
    """Parity AT the benchmarked training shape (BASELINE configs[2]: ACDC, batch 24): train-mode BatchNorm couples the whole
    batch, and the oracle's autograd at batch 24 would need > 60 GB of host memory (it materialises every N x N map), so the
    bf16 product step is checked against the fp32 validation engine -- itself pinned to the oracle / reference goldens at batch
    2 -- on identical weights, inputs and labels: loss, logits, direction and length of the full gradient."""
    kw = fixtures.CONFIGS["acdc"]
    from cenet_b200.networks import CENet
    torch.manual_seed(1234)
    sd = fixtures.perturb_state(CENet(**kw).state_dict(), 1234)
    x = fixtures.synth_input("acdc", 24).to(DEV)
    labels = torch.randint(0, 4, (24, 224, 224), generator=torch.Generator().manual_seed(5)).to(DEV)
    res = {}
    for prec in ("fp32", "bf16"):
        m, eng = _engine("acdc", prec, sd)
        eng.use_graph = False
        out = eng.train_step(x, labels, optimize=False)
        torch.cuda.synchronize()
        res[prec] = (out[0].item(), eng.buf("logits", (24, 4, 224, 224), torch.float32).clone(), eng.gflat.clone())
        del m, eng
        torch.cuda.empty_cache()
    (l32, y32, g32), (l16, y16, g16) = res["fp32"], res["bf16"]
    assert abs(l16 - l32) < 1e-2 * max(1.0, abs(l32)), (l16, l32)
    e = ((y16 - y32).norm() / y32.norm()).item()
    assert e < 2e-2, e
    cos = (torch.dot(g16, g32) / (g16.norm() * g32.norm())).item()
    assert cos > 0.99 and abs(g16.norm().item() / g32.norm().item() - 1.0) < 0.05, (cos, g16.norm().item(), g32.norm().item())


# ---------------------------------------------------------------------------------------------- pinned to the REAL reference
@pytest.mark.parametrize("fx", ["train_acdc_b2_s224", "train_synapse_b2_s96", "train_acdc_b1_s64"])
def test_train_step_fp32_matches_reference_train_mode_golden(fx):
    """One fp32 step on the B200 vs tests/golden/train_*.pt = the reference module itself in train() mode with its own
    Criterion('dice,ce'): loss, logits, the norm of all 630 gradients, sampled gradients across every module family, and
    every BatchNorm running_mean / running_var / num_batches_tracked after the step (incl. the CCU B>1 guard at B=1)."""
    from test_oracle_golden import check_train_against_golden, rebuild_train_case
    from conftest import GOLDEN
    g = torch.load(os.path.join(GOLDEN, fx + ".pt"), weights_only=False)
    sd, kw, x, labels = rebuild_train_case(g)
    m, eng = _engine(g["config"], "fp32", sd)
    eng.use_graph = False
    out = eng.train_step(x.to(DEV), labels.to(DEV), optimize=False)
    torch.cuda.synchronize()
    logits = eng.buf("logits", (g["batch"], kw["num_classes"], g["size"], g["size"]), torch.float32).cpu()
    grads = {k: v.cpu() for k, v in eng.GP.items()}
    buffers = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    check_train_against_golden(g, out[0].item(), logits, grads, buffers, 1e-4, fx)


def test_eval_after_fused_steps_uses_new_weights():
    """TrainEngine.train_step rewrites parameters and BatchNorm statistics through raw pointers (no tensor
    `_version` bump).  An eval forward that already packed its weights must re-pack after it: eval -> N fused steps -> eval
    equals a freshly built module holding the trained state_dict."""
    from cenet_b200.networks import CENet
    kw, sd, x, labels, *_ = _ref("acdc", 2, 224)
    m, eng = _engine("acdc", "bf16", sd)
    xd, ld = x.to(DEV), labels.to(DEV)
    m.eval()
    with torch.no_grad():
        y0 = m(xd).clone()                                        # packs the inference weights + captures the graph
    m.train()
    for _ in range(3):
        eng.train_step(xd, ld, lr=1e-3)
    m.eval()
    with torch.no_grad():
        y1 = m(xd).clone()
    fresh = CENet(**kw)
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
    fresh = fresh.to(DEV).eval()
    with torch.no_grad():
        y2 = fresh(xd)
    assert (y1 - y0).abs().max().item() > 1e-3                    # training moved the output ...
    assert torch.equal(y1, y2)                                    # ... and the cached engine followed it exactly


@pytest.mark.parametrize("name,batch,size", [("acdc", 2, 96), ("skin", 1, 96)])
def test_eval_mode_backward_matches_oracle_autograd(name, batch, size, monkeypatch):
    """SURVEY 8b "Call": the autograd path must also work in eval() (running-statistics BatchNorm, no DropPath; B = 1 included:
    the CCU skips its BatchNorm1d).  `net.eval(); loss = crit(net(x)); loss.backward()` on the drop-in (fp32 validation precision)
    against torch autograd through the oracle's eval-mode forward: every parameter gradient, and the buffers must not move."""
    from cenet_b200.networks import CENet
    monkeypatch.setenv("CENET_B200_PRECISION", "fp32")
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    sd = fixtures.perturb_state(CENet(**kw).state_dict(), 1234)
    x = fixtures.synth_input(name, batch, size=size)
    labels = torch.randint(0, kw["num_classes"], (batch, size, size), generator=torch.Generator().manual_seed(5))
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits_ref = O.cenet_forward(leaf, O.Cfg(**kw), x, training=False)
    loss_ref = O.criterion_dice_ce(logits_ref, labels, kw["num_classes"])
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [leaf[k] for k in names], allow_unused=True)))
    m = CENet(**kw)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    y = m(x.to(DEV))                                                      # eval() forward under grad mode
    assert y.requires_grad
    assert ((y.detach().cpu() - logits_ref.detach()).norm() / logits_ref.norm()).item() < 1e-4
    yl = y.detach().cpu().requires_grad_(True)
    loss = O.criterion_dice_ce(yl, labels, kw["num_classes"])
    loss.backward()
    y.backward(yl.grad.to(DEV))
    bad = []
    gn = max(g.norm().item() for g in gref.values() if g is not None)
    for k, p in m.named_parameters():
        g = gref[k]
        if g is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0
            continue
        err = (p.grad.cpu() - g).norm().item()
        if not err < 2e-3 * g.norm().item() + 1e-6 * gn:
            bad.append((k, err / max(g.norm().item(), 1e-12)))
    assert not bad, bad[:20]
    after = m.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            assert torch.equal(after[k].cpu(), v), k                   # eval mode: statistics untouched


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_step_with_droppath_matches_oracle_given_the_same_masks(precision):
    """DropPath ON, as the bench runs it: the masks the step drew are read back (TrainEngine.dp_scale) and handed to the oracle.
    fp32: loss + every gradient to 2e-3.  bf16: the product path, where the tcgen05 weight gradient runs in MASK mode (dropped
    samples skipped, one common scale) -- gradient cosine / norm like the DropPath-free bf16 test."""
    from cenet_b200.networks import CENet
    from cenet_b200.train import TrainEngine
    kw = fixtures.CONFIGS["acdc"]
    torch.manual_seed(1234)
    m = CENet(**kw)
    sd = fixtures.perturb_state(m.state_dict(), 1234)
    m.load_state_dict(sd)
    m.backbone.drop_path_probs = [min(0.6, 4.0 * p) for p in m.backbone.drop_path_probs]
    m = m.to(DEV).train()
    eng = TrainEngine(m, DEV, precision)
    eng.use_graph = False
    batch, size = 4, 96
    x = fixtures.synth_input("acdc", batch, size=size)
    labels = torch.randint(0, 4, (batch, size, size), generator=torch.Generator().manual_seed(5))
    torch.manual_seed(11)
    out = eng.train_step(x.to(DEV), labels.to(DEV), optimize=False)
    torch.cuda.synchronize()
    masks = eng.dp_scale.float().cpu().clone()
    assert (masks == 0).any() and (masks > 1).any()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True, drop_masks=masks)
    loss_ref = O.criterion_dice_ce(logits, labels, 4)
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [leaf[k] for k in names], allow_unused=True)))
    if precision == "fp32":
        assert abs(out[0].item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
        gn = max(g.norm().item() for g in gref.values() if g is not None)
        bad = [(k, (eng.GP[k].cpu() - g).norm().item() / max(g.norm().item(), 1e-12)) for k, g in gref.items()
               if g is not None and not (eng.GP[k].cpu() - g).norm().item() < 2e-3 * g.norm().item() + 1e-6 * gn]
        assert not bad, bad[:20]
    else:
        assert abs(out[0].item() - loss_ref.item()) < 1e-2 * max(1.0, abs(loss_ref.item()))
        a = torch.cat([eng.GP[k].cpu().flatten() for k, g in gref.items() if g is not None])
        b = torch.cat([g.flatten() for g in gref.values() if g is not None])
        cos = torch.dot(a, b) / (a.norm() * b.norm())
        assert cos > 0.99 and abs(a.norm() / b.norm() - 1) < 0.05, (cos.item(), (a.norm() / b.norm()).item())
