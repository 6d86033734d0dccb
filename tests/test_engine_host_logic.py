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


@pytest.mark.parametrize("name,batch,flash", [("acdc", 1, False), ("synapse", 2, True), ("skin", 1, True)])
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
