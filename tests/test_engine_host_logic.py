"""Host logic of the launch plan on CPU: `Engine` is run with tests/fake_ops.py (a torch emulation of the C-ABI
ops, same pitches / offsets / epilogue semantics) in place of the CUDA library and must reproduce the oracle.
This checks weight folding (BatchNorm, layer scale, non-local mix), buffer wiring and slice arithmetic; the kernels
themselves are checked on the B200 by tests/test_gpu_*.py."""
import pytest
import torch

import fake_ops
from oracle import cenet_oracle as O
from oracle import fixtures


def _engine(name, precision="fp32", flash=False):
    import cenet_b200.engine as E
    from cenet_b200.networks import CENet
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(1234)
    m = CENet(**kw)
    sd = fixtures.perturb_state(m.state_dict(), 1234)
    m.load_state_dict(sd)
    m.eval()
    eng = E.Engine(m, "cpu", precision)
    eng.use_graph = False
    eng.use_flash = flash
    return eng, sd, kw


@pytest.fixture(autouse=True)
def _patch_ops(monkeypatch):
    import cenet_b200.engine as E
    monkeypatch.setattr(E, "ops", fake_ops)


@pytest.mark.parametrize("name,batch,flash", [("acdc", 1, False), ("synapse", 2, True), ("skin", 1, True),
                                              ("acdc_b1", 1, True), ("acdc_add", 1, True), ("synapse_uprb", 1, True),
                                              ("acdc_uptc", 1, True)])
def test_launch_plan_reproduces_oracle(name, batch, flash):
    eng, sd, kw = _engine(name, flash=flash)
    x = fixtures.synth_input(name, batch)
    taps_ref = {}
    with torch.no_grad():
        y_ref = O.cenet_forward(sd, O.Cfg(**kw), x, taps=taps_ref)
    eng.taps = {}
    y = eng.forward(x)
    for k, v in eng.taps.items():
        e = ((v - taps_ref[k]).norm() / taps_ref[k].norm()).item()
        assert e < 2e-5, (k, e)
    e = ((y - y_ref).norm() / y_ref.norm()).item()
    assert e < 2e-5, e
    lab = eng.forward(x, labels=True)
    assert torch.equal(lab, O.predict_labels(y))


def test_repack_on_weight_change_and_batch_keys():
    eng, sd, kw = _engine("acdc")
    x = fixtures.synth_input("acdc", 1, size=64)            # any multiple of 32 works on the host side
    y0 = eng.forward(x).clone()
    with torch.no_grad():
        eng.mod.out.out[1].conv.conv.bias.add_(2.0)
    y1 = eng.forward(x)
    torch.testing.assert_close(y1, y0 + 2.0, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 1, 100, 100))
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 3, 64, 64))


# ---------------------------------------------------------------------------------------------- CENetOrg
def _org_engine(precision="fp32", flash=True):
    import os
    import cenet_b200.engine as E
    import cenet_b200.engine_org as EO
    from cenet_b200.networks import CENetOrg
    from conftest import GOLDEN
    g = torch.load(os.path.join(GOLDEN, "model_org_synapse_b2.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    m = CENetOrg(**g["kw"])
    sd = fixtures.perturb_state(m.state_dict(), g["seed"])
    sd["out.conv.conv.weight"] = sd["out.conv.conv.weight"] * 20.0
    m.load_state_dict(sd)
    m.eval()
    eng = EO.EngineOrg(m, "cpu", precision)
    eng.use_graph = False
    eng.use_flash = flash
    return m, eng, g


def test_cenet_org_launch_plan_reproduces_the_reference(monkeypatch):
    """The CENetOrg launch plan on the torch emulation of the kernels vs the REAL reference's outputs (tests/golden/
    model_org_synapse_b2.pt: cenet_org.net.Net with the TEST_ORG kwargs of scripts/synapse.sh): 822-key state_dict,
    strided logits, every decoder tap and the integer label map."""
    import cenet_b200.engine_org as EO
    monkeypatch.setattr(EO, "ops", fake_ops)
    m, eng, g = _org_engine()
    assert len(m.state_dict()) == g["n_keys"] == 822
    x = fixtures.synth_input("synapse", g["batch"], seed=g["input_seed"])
    eng.taps = {}
    y = eng.forward(x)
    for k, ref in g["taps"].items():
        v = eng.taps[k]
        s = v.flatten()[:: max(1, v.numel() // 4096)][:4096]
        e = ((s - ref["sample"]).norm() / ref["sample"].norm()).item()
        assert e < 5e-5, (k, e)
        assert abs(v.norm().item() - ref["norm"]) < 1e-4 * ref["norm"], k
    e = ((y[:, :, ::4, ::4] - g["logits_strided"]).norm() / g["logits_strided"].norm()).item()
    assert e < 5e-5, e
    lab = eng.forward(x, labels=True)
    assert torch.equal(lab, O.predict_labels(y))
    assert (lab[:, ::2, ::2] == g["labels_strided"]).float().mean().item() > 0.9999


def test_cenet_org_contract():
    from cenet_b200.networks import CENetOrg
    m = CENetOrg(num_classes=9, input_channels=1, scale_factors=[0.8, 0.4], encoder="pvt_v2_b2", pretrain=False, num_heads=[16, 8, 8])
    assert len(m.state_dict()) == 822 and sum(p.numel() for p in m.parameters()) == 33379849
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 1, 224, 224))                    # no CPU path
    with pytest.raises(NotImplementedError):
        CENetOrg(encoder="resnet50")


def test_bf16_plan_variants_on_the_emulation():
    """the bf16 launch plan (implicit strided convolutions through conv_nhwc, flash attention entry points) and its switchable
    variants -- fused Mix-FFN tail on, im2col strided convolutions -- agree with each other and with the oracle on the emulation"""
    x = fixtures.synth_input("synapse", 2, size=64)
    outs = {}
    for tag, attrs in (("default", {}), ("mixffn", {"fuse_mixffn": True}), ("im2col", {"implicit_strided": False})):
        eng, sd, kw = _engine("synapse", precision="bf16", flash=True)
        assert eng.implicit_strided and not eng.fuse_mixffn
        for k, v in attrs.items():
            setattr(eng, k, v)
        n0 = fake_ops.launch_count()
        outs[tag] = eng.forward(x).float().clone()
        outs[tag + ".launches"] = fake_ops.launch_count() - n0
    with torch.no_grad():
        y_ref = O.cenet_forward(sd, O.Cfg(**kw), x)
    for tag in ("default", "mixffn", "im2col"):
        e = ((outs[tag] - y_ref).norm() / y_ref.norm()).item()
        assert e < 2e-2, (tag, e)
        assert ((outs[tag] - outs["default"]).norm() / outs["default"].norm()).item() < 1e-2, tag
    assert outs["mixffn.launches"] == outs["default.launches"] - 7          # dwconv3x3 + fc2 -> one launch in the 3 + 4 blocks of stages 1-2
    assert outs["im2col.launches"] > outs["default.launches"]
