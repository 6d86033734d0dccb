"""Pins the CPU oracle (oracle/cenet_oracle.py) to outputs of the REAL reference recorded in tests/golden/ by
tests/golden/make_golden.py.  Runs on CPU; this is what makes the oracle trustworthy as the GPU tests' checker."""
import os

import pytest
import torch

from oracle import cenet_oracle as O
from oracle import fixtures

from conftest import GOLDEN

TOL = dict(rtol=1e-5, atol=1e-5)


def _case(golden_modules, key):
    c = golden_modules[key]
    return c["state"], c["inputs"], c["output"]


def test_diffattn(golden_modules):
    for key, heads, depth in (("diffattn_e64_h2_n80", 2, 2), ("diffattn_e32_h2_n64", 2, 3)):
        sd, (x,), y = _case(golden_modules, key)
        sd = {"m." + k: v for k, v in sd.items()}
        torch.testing.assert_close(O.diff_attention(sd, "m", x, heads, depth), y, **TOL)


def test_fea(golden_modules):
    for key, sc in (("fea_2scales", [0.8, 0.4]), ("fea_3scales", [1.0, 0.75, 0.5])):
        sd, (x,), y = _case(golden_modules, key)
        torch.testing.assert_close(O.fea({"m." + k: v for k, v in sd.items()}, "m", x, sc), y, **TOL)


def test_dseb(golden_modules):
    sd, (skip, dec), y = _case(golden_modules, "dseb_c16")
    out = O.dse_block({"m." + k: v for k, v in sd.items()}, "m", skip, dec, [0.8, 0.4], 2, 3)
    torch.testing.assert_close(out, y, **TOL)


def test_nonlocal_ccu_srm(golden_modules):
    sd, (x,), y = _case(golden_modules, "nonlocal_c64")
    torch.testing.assert_close(O.nonlocal_block({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)
    for key in ("ccu_c64_b2", "ccu_c64_b1"):                       # B == 1 skips BatchNorm1d (cfam.py:260-261)
        sd, (x,), y = _case(golden_modules, key)
        torch.testing.assert_close(O.ccu({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)
    sd, (x,), y = _case(golden_modules, "srm")
    torch.testing.assert_close(O.srm({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)


def test_cfam_and_multi_order(golden_modules):
    sd, (x,), y = _case(golden_modules, "modw_c64")
    torch.testing.assert_close(O.multi_order_dwconv({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)
    sd, (x,), y = _case(golden_modules, "cfam_c64")
    torch.testing.assert_close(O.cfa_module({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)


def test_up_blocks(golden_modules):
    sd, (x,), y = _case(golden_modules, "eucb")
    torch.testing.assert_close(O.eucb({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)
    sd, (x,), y = _case(golden_modules, "upconv")
    torch.testing.assert_close(O.up_conv({"m." + k: v for k, v in sd.items()}, "m", x), y, **TOL)


def test_encoder_pieces(golden_modules):
    sd, (x, H, W), y = _case(golden_modules, "pvt_attn_sr2")
    torch.testing.assert_close(O.sr_attention({"m." + k: v for k, v in sd.items()}, "m", x, H, W, 2, 2), y, **TOL)
    sd, (x, H, W), y = _case(golden_modules, "pvt_block_sr1")
    torch.testing.assert_close(O.pvt_block({"m." + k: v for k, v in sd.items()}, "m", x, H, W, 1, 1), y, **TOL)


def test_out_head(golden_modules):
    sd, (dec, x), y = _case(golden_modules, "outhead")
    cfg = O.Cfg(input_channels=1, num_classes=4, out_up_block="upcn")
    torch.testing.assert_close(O.out_head({"out." + k: v for k, v in sd.items()}, cfg, dec, x), y, **TOL)


def test_loss(golden_loss):
    for key, c in golden_loss.items():
        ncls = int(key[1:])
        logits = c["logits"].clone().requires_grad_(True)
        loss = O.criterion_dice_ce(logits, c["labels"], ncls)
        loss.backward()
        torch.testing.assert_close(loss.detach(), c["loss"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logits.grad, c["grad"], rtol=1e-4, atol=1e-8)


def _weights(c):
    w = dict(zip(c["loss_type"].split(","), (float(v) for v in c["loss_weights"].split(","))))
    return dict(w_dice=w.get("dice", 0.0), w_ce=w.get("ce", 0.0), w_boundary=w.get("boundary", 0.0))


def test_boundary_dou_loss(golden_loss_boundary):
    """BoundaryDoULoss alone and inside Criterion('dice,ce,boundary') as the reference computes them (core.py:83-131,161-188)"""
    for key, c in golden_loss_boundary.items():
        ncls = c["logits"].shape[1]
        logits = c["logits"].clone().requires_grad_(True)
        loss = O.criterion(logits, c["labels"], ncls, **_weights(c))
        loss.backward()
        torch.testing.assert_close(loss.detach(), c["loss"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logits.grad, c["grad"], rtol=1e-4, atol=1e-8)
        C, S = O.boundary_counts(c["labels"], ncls)
        assert torch.equal(S, torch.bincount(c["labels"].flatten(), minlength=ncls)) and bool((C <= S).all())


@pytest.mark.parametrize("name,batch", [("acdc", 1), ("synapse", 2), ("skin", 1), ("acdc_b1", 1), ("acdc_b5", 1),
                                        ("acdc_add", 1), ("synapse_uprb", 1), ("acdc_uptc", 1)])
def test_whole_model(name, batch):
    """Rebuild the deterministic weights, run the oracle, compare with what the reference produced."""
    from cenet_b200.networks import CENet
    g = torch.load(os.path.join(GOLDEN, f"model_{name}_b{batch}.pt"), weights_only=False)
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(g["seed"])
    sd = fixtures.perturb_state(CENet(**kw).state_dict(), g["seed"])
    x = fixtures.synth_input(name, batch, seed=g["input_seed"])
    taps = {}
    with torch.no_grad():
        y = O.cenet_forward(sd, O.Cfg(**kw), x, taps=taps)
    torch.testing.assert_close(y[:, :, ::8, ::8], g["logits_strided"], rtol=1e-4, atol=1e-5)
    assert abs(y.std().item() - g["logits_std"]) < 1e-4
    lab = O.predict_labels(y)
    assert torch.equal(torch.bincount(lab.flatten(), minlength=kw["num_classes"]), g["label_hist"])
    assert torch.equal(lab[:, ::4, ::4], g["labels_strided"])
    for k, ref in g["taps"].items():
        v = taps[k]
        samp = v.flatten()[:: max(1, v.numel() // 512)][:512]
        torch.testing.assert_close(samp, ref["sample"], rtol=1e-4, atol=1e-5, msg=lambda m, k=k: f"tap {k}: {m}")


# ---------------------------------------------------------------------------------------------- train mode
def _train_golden(fx):
    return torch.load(os.path.join(GOLDEN, fx + ".pt"), weights_only=False)


def rebuild_train_case(g):
    """state dict / input / labels of a tests/golden/train_*.pt fixture (same recipe as make_golden.train_fixture)"""
    from cenet_b200.networks import CENet
    kw = fixtures.CONFIGS[g["config"]]
    torch.manual_seed(g["seed"])
    sd = fixtures.perturb_state(CENet(**kw).state_dict(), g["seed"])
    x = fixtures.synth_input(g["config"], g["batch"], size=g["size"], seed=g["input_seed"])
    labels = torch.randint(0, kw["num_classes"], (g["batch"], g["size"], g["size"]),
                           generator=torch.Generator().manual_seed(g["label_seed"]))
    return sd, kw, x, labels


def check_train_against_golden(g, loss, logits, grads, buffers, rtol, what):
    """shared by the CPU oracle test below and the GPU test (tests/test_gpu_train_model.py): loss, strided logits, the norm
    of EVERY parameter gradient, sampled gradients of 51 tensors across all module families, and every BatchNorm buffer
    (running_mean / running_var / num_batches_tracked) after one train-mode step, against the REAL reference."""
    assert abs(float(loss) - float(g["loss"])) < rtol * max(1.0, abs(float(g["loss"]))), (what, float(loss), float(g["loss"]))
    ls = logits[:, :, ::4, ::4]
    e = ((ls - g["logits_strided"]).norm() / g["logits_strided"].norm()).item()
    assert e < rtol, (what, "logits", e)
    bad = []
    gmax = max(v for v in g["grad_norms"].values() if v is not None)
    for n, ref_norm in g["grad_norms"].items():
        if ref_norm is None:
            continue
        mine = grads[n].float().norm().item()
        if abs(mine - ref_norm) > rtol * 20 * max(ref_norm, 1e-4 * gmax):
            bad.append((n, mine, ref_norm))
    assert not bad, (what, "grad norms", bad[:8], len(bad))
    for n, ref_s in g["grad_samples"].items():
        f = grads[n].float().flatten().cpu()
        mine = f[:: max(1, f.numel() // 2048)][:2048]
        e = ((mine - ref_s).norm() / max(ref_s.norm().item(), 1e-4 * gmax)).item()
        if e > rtol * 20:
            bad.append((n, e))
    assert not bad, (what, "grad samples", bad[:8], len(bad))
    for k, v in g["buffers_after"].items():
        if k.endswith("num_batches_tracked"):
            if int(buffers[k]) != int(v):
                bad.append((k, int(buffers[k]), int(v)))
        else:
            e = ((buffers[k].float().cpu() - v).norm() / v.norm()).item()
            if e > rtol:
                bad.append((k, e))
    assert not bad, (what, "BatchNorm buffers", bad[:8], len(bad))


@pytest.mark.parametrize("fx", ["train_acdc_b2_s224", "train_synapse_b2_s96", "train_acdc_b1_s64"])
def test_oracle_train_mode_pinned_to_reference(fx):
    """oracle.cenet_forward(training=True) + autograd + criterion_dice_ce == the reference module in train() mode with its
    own utils/core.py Criterion('dice,ce'): loss, logits, all 630 gradient norms, gradient samples, BN buffers after."""
    g = _train_golden(fx)
    sd, kw, x, labels = rebuild_train_case(g)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in sd.items()}
    stats = {}
    logits = O.cenet_forward(leaf, O.Cfg(**kw), x, training=True, new_stats=stats)
    loss = O.criterion_dice_ce(logits, labels, kw["num_classes"])
    gr = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    grads = {k: (v if v is not None else torch.zeros_like(sd[k])) for k, v in zip(names, gr)}
    buffers = dict(sd)
    buffers.update(stats)
    if g["batch"] == 1:                                             # CCU's BatchNorm1d never ran (cfam.py:260-261)
        assert not any(".ccu.bn." in k for k in stats)
    check_train_against_golden(g, loss.item(), logits.detach(), grads, buffers, 2e-5, fx)
