"""Host logic of the TRAINING launch plan on CPU: `TrainEngine` runs with tests/fake_ops.py + tests/fake_train_ops.py
(torch emulations of the C-ABI kernels) and must reproduce loss and EVERY parameter gradient of autograd through the
oracle in train mode (batch-statistics BatchNorm).  This pins the tape order, gradient accumulation across fan-outs,
channel-slice bookkeeping, weight (re)packing and the flat parameter / gradient buffers; the CUDA kernels themselves
are checked against the same emulations on the B200 (tests/test_gpu_train_ops.py)."""
import pytest
import torch

import fake_ops
import fake_train_ops
from oracle import cenet_oracle as O
from oracle import fixtures


@pytest.fixture(autouse=True)
def _patch_ops(monkeypatch):
    import cenet_b200.train as T
    monkeypatch.setattr(T, "ops", fake_ops)
    monkeypatch.setattr(T, "tops", fake_train_ops)


def _build(name, flash=False):
    import cenet_b200.train as T
    from cenet_b200.networks import CENet
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    m = CENet(**kw)
    sd = fixtures.perturb_state(m.state_dict(), 1234)
    m.load_state_dict(sd)
    m.train()
    eng = T.TrainEngine(m, "cpu", "fp32")
    eng.use_graph = False
    eng.use_flash = flash
    eng.drop_path = False
    return m, eng, sd, kw


def _oracle_grads(sd, kw, x, labels):
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True)
    loss = O.criterion_dice_ce(logits, labels, kw["num_classes"])
    grads = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    return loss.item(), logits.detach(), dict(zip(names, grads))


@pytest.mark.parametrize("name,batch,size,flash", [("acdc", 2, 64, False), ("synapse", 2, 64, True), ("skin", 2, 64, True),
                                                   ("acdc", 1, 224, True), ("acdc_add", 2, 64, True),
                                                   ("synapse_uprb", 2, 64, True), ("acdc_b1", 2, 64, True), ("acdc_uptc", 2, 64, True)])
def test_train_step_matches_oracle_autograd(name, batch, size, flash):
    m, eng, sd, kw = _build(name, flash)
    x = fixtures.synth_input(name, batch, size=size)
    g = torch.Generator().manual_seed(5)
    labels = torch.randint(0, kw["num_classes"], (batch, size, size), generator=g)
    loss_ref, logits_ref, gref = _oracle_grads(sd, kw, x, labels)
    out = eng.train_step(x, labels, optimize=False)
    assert abs(out[0].item() - loss_ref) < 2e-5 * max(1.0, abs(loss_ref)), (out[0].item(), loss_ref)
    bad = []
    for k, gr in gref.items():
        mine = eng.GP[k]
        if gr is None:                                     # CCU BatchNorm1d is skipped at B == 1 (cfam.py:260)
            assert mine.abs().max().item() == 0.0, k
            continue
        # (biases feeding a BatchNorm / softmax have analytically zero gradients: absolute floor)
        err = (mine - gr).norm().item()
        if not err < 2e-3 * gr.norm().item() + 1e-6:
            bad.append((k, err, gr.norm().item()))
    assert not bad, bad[:20]


@pytest.mark.parametrize("name,batch", [("acdc", 2), ("skin", 1)])
def test_frozen_statistics_pass_matches_oracle_eval_autograd(name, batch):
    """eval() semantics inside a gradient pass (what `net.eval(); net(x).backward()` runs, networks.cenet._EvalForward): BatchNorm,
    the CCU's BatchNorm1d and the SRM's BatchNorm2d(1) use their running statistics and update nothing; loss and every parameter
    gradient against autograd through the oracle's EVAL-mode forward."""
    m, eng, sd, kw = _build(name, True)
    size = 64
    x = fixtures.synth_input(name, batch, size=size)
    labels = torch.randint(0, kw["num_classes"], (batch, size, size), generator=torch.Generator().manual_seed(5))
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=False)
    loss_ref = O.criterion_dice_ce(logits, labels, kw["num_classes"])
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [leaf[k] for k in names], allow_unused=True)))
    eng.frozen_stats = True
    out = eng.train_step(x, labels, optimize=False)
    assert abs(out[0].item() - loss_ref.item()) < 2e-5 * max(1.0, abs(loss_ref.item()))
    bad = []
    for k, gr in gref.items():
        mine = eng.GP[k]
        if gr is None:
            assert mine.abs().max().item() == 0.0, k
            continue
        err = (mine - gr).norm().item()
        if not err < 2e-3 * gr.norm().item() + 1e-6:
            bad.append((k, err, gr.norm().item()))
    assert not bad, bad[:20]
    after = m.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            assert torch.equal(after[k], v), k


def test_train_step_with_droppath_matches_oracle_given_the_same_masks():
    """DropPath ON (as in the bench): the step's own Bernoulli draws (two independent rows per block, pvtv2.py:146-147) are read
    back and handed to the oracle; loss and every gradient must then agree -- this also drives the mask mode of the weight
    gradient (dropped samples skipped) through the whole model."""
    m, eng, sd, kw = _build("acdc", True)
    eng.drop_path = True
    m.backbone.drop_path_probs = [min(0.6, 4.0 * p) for p in m.backbone.drop_path_probs]      # enough drops at batch 4
    eng.dp_keep = (1.0 - torch.tensor(m.backbone.drop_path_probs).repeat_interleave(2).view(-1, 1))
    batch, size = 4, 64
    x = fixtures.synth_input("acdc", batch, size=size)
    labels = torch.randint(0, 4, (batch, size, size), generator=torch.Generator().manual_seed(5))
    torch.manual_seed(11)
    out = eng.train_step(x, labels, optimize=False)
    masks = eng.dp_scale.clone()
    assert (masks == 0).any() and (masks > 1).any()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True, drop_masks=masks)
    loss_ref = O.criterion_dice_ce(logits, labels, 4)
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [leaf[k] for k in names], allow_unused=True)))
    assert abs(out[0].item() - loss_ref.item()) < 2e-5 * max(1.0, abs(loss_ref.item()))
    bad = []
    for k, gr in gref.items():
        if gr is None:
            continue
        err = (eng.GP[k] - gr).norm().item()
        if not err < 2e-3 * gr.norm().item() + 1e-6:
            bad.append((k, err, gr.norm().item()))
    assert not bad, bad[:20]


def test_running_stats_and_adamw_step():
    m, eng, sd, kw = _build("acdc")
    x = fixtures.synth_input("acdc", 2, size=64)
    labels = torch.randint(0, 4, (2, 64, 64), generator=torch.Generator().manual_seed(5))
    p0 = eng.pflat.clone()
    out = eng.train_step(x, labels, lr=1e-3, weight_decay=1e-2)
    g = eng.gflat.clone()
    # torch.optim.AdamW, first step: p <- p(1 - lr wd) - lr * g / (|g| + eps)
    ref = p0 * (1 - 1e-3 * 1e-2) - 1e-3 * g / (g.abs() + 1e-8)
    assert (eng.pflat - ref).abs().max().item() < 1e-6
    # module parameters ARE the flat buffer
    for n, p in m.named_parameters():
        assert p.data_ptr() == eng.P[n].data_ptr()
    # BatchNorm bookkeeping: momentum 0.1 towards the batch statistics, counters incremented
    assert int(m.state_dict()["decoder.dec1.norm1.num_batches_tracked"]) == 8
    rm = m.state_dict()["out.out.0.norm1.running_mean"]
    assert not torch.allclose(rm, sd["out.out.0.norm1.running_mean"])
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("name", ["acdc", "skin"])
def test_gathered_weight_pack_equals_torch_layouts(name):
    """pack(): the one-launch index-map re-pack must reproduce every torch-evaluated layout, also after a weight update."""
    m, eng, sd, kw = _build(name)
    eng.pack()                                            # CPU engines keep the torch path ...
    assert eng._pack_maps is None
    eng._build_pack_maps()                                # ... so build (and self-check) the maps explicitly
    assert eng._pack_maps and sum(i.numel() for i, _ in eng._pack_maps) > 60e6
    with torch.no_grad():
        eng.pflat.mul_(1.5).add_(0.01)                    # what an optimizer step does
    eng.pack()                                            # gather path
    got = {k: v.clone() for k, v in eng.w.items()}
    eng._pack_maps = None
    eng.w = {}
    eng._pack_torch()
    assert set(got) == set(eng.w)
    for k, v in eng.w.items():
        assert torch.equal(got[k], v), k
