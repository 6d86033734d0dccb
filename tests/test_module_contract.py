"""Drop-in boundary: constructor signature, state_dict contract, host-side behaviour.  CPU-only."""
import copy
import os
import inspect

import pytest
import torch

from cenet_b200.networks import CENet, CENetOrg
from oracle import fixtures


def test_constructor_signature_matches_reference():
    sig = inspect.signature(CENet.__init__)
    names = list(sig.parameters)[1:]
    assert names == ["input_channels", "num_classes", "scale_factors", "diffatt_num_heads", "encoder", "enc_pretrain",
                     "freeze_bb", "skip_mode", "dec_up_block", "out_merge_mode", "out_up_block", "out_up_ks", "writer",
                     "base_ptdir"]                                      # net.py:9-22
    d = {k: v.default for k, v in sig.parameters.items() if k != "self"}
    assert d["input_channels"] == 1 and d["num_classes"] == 1 and d["scale_factors"] == [0.8, 0.4]
    assert d["diffatt_num_heads"] == [2, 2, 2] and d["dec_up_block"] == "eucb" and d["out_up_block"] == "eucb"


@pytest.mark.parametrize("name,nparams", [("acdc", 33384872), ("synapse", 33384861), ("skin", 33386918)])
def test_state_dict_contract(name, nparams):
    m = CENet(**fixtures.CONFIGS[name])
    sd = m.state_dict()
    assert len(sd) == 801                                              # 630 params + 171 buffers
    assert sum(p.numel() for p in m.parameters()) == nparams
    assert len(list(m.parameters())) == 630
    for k, shape in {"backbone.block1.0.attn.sr.weight": (64, 64, 8, 8),
                     "decoder.dec1.mca.value.dlps.3.1.weight": (4, 4, 1, 1),
                     "out.rb.0.conv2.conv.weight": (32, 32, 5, 5),
                     "decoder.skip_enhancer1.boundary.w": (1, 128, 1, 1),
                     "decoder.dec4.mca.ccu.fc1.weight": (1536, 1, 3)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "out.out.1.conv.conv.bias" in sd and "decoder.dec2.mca.denoising_module.w" in sd
    # round trip + attribute access used by the reference mains (utils.py:179-180)
    m2 = CENet(**fixtures.CONFIGS[name])
    m2.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in m.backbone.parameters()) > 0 and sum(p.numel() for p in m.decoder.parameters()) > 0


def test_module_census_and_init_statistics():
    torch.manual_seed(0)
    m = CENet(**fixtures.CONFIGS["acdc"])
    import collections
    c = collections.Counter(type(x).__name__ for x in m.modules())
    assert (c["Conv2d"], c["Linear"], c["Conv1d"], c["LayerNorm"], c["BatchNorm2d"], c["BatchNorm1d"]) == \
        (125, 92, 8, 53, 53, 4)
    sd = m.state_dict()
    assert abs(sd["backbone.block1.0.attn.q.weight"].std().item() - 0.02) < 2e-3          # trunc_normal(.02)
    assert abs(sd["backbone.block1.0.attn.sr.weight"].std().item() - (2.0 / (8 * 8 * 64)) ** 0.5) < 2e-3
    assert abs(sd["decoder.up3.up_dwc.1.weight"].std().item() - 0.02) < 3e-3              # 'normal' scheme
    assert torch.all(sd["decoder.dec1.layer_scale_1"] == 1e-6)
    assert abs(sd["decoder.skip_enhancer1.boundary.w"].mean().item() - 0.5) < 0.3
    assert sd["decoder.dec1.mca.denoising_module.w"].item() == 0.5


def test_deepcopy_and_modes():
    m = CENet(**fixtures.CONFIGS["acdc"])
    m2 = copy.deepcopy(m)                                              # utils.py:113 (thop deep-copies the model)
    assert m2.state_dict().keys() == m.state_dict().keys()
    m.eval(); m.train()


def test_error_conventions():
    with pytest.raises(AssertionError):
        CENet(dec_up_block="bogus")                                    # decoders.py:47
    with pytest.raises(AssertionError):
        CENet(out_merge_mode="bogus")                                  # out.py:31
    CENetOrg()                                                         # f2: built (inference path)
    with pytest.raises(NotImplementedError):
        CENetOrg(skip_mode="add")
    CENet(encoder="not_an_encoder")                                    # encoder.py:48-52: silent fallback to b2
    with pytest.raises(NotImplementedError):
        CENet(encoder="pvt_v2_b0")                                     # widths 32..256: not built
    for kind in ("uprb", "uptc", "upcn", "eucb"):                      # decoders.py:47-60: all four up blocks construct
        CENet(dec_up_block=kind, out_up_block=kind)
    CENet(skip_mode="add", out_merge_mode="add")
    n1 = sum(p.numel() for p in CENet(encoder="pvt_v2_b1").backbone.parameters())
    n3 = sum(p.numel() for p in CENet(encoder="pvt_v2_b3").backbone.parameters())
    assert n1 < n3                                                     # pvtv2.py:392-413: depths (2,2,2,2) vs (3,4,18,3)


def test_no_cpu_path():
    m = CENet(**fixtures.CONFIGS["acdc"]).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 224, 224))


def test_run_main_launcher_resolves_networks_to_the_drop_in(tmp_path):
    """`python -m cenet_b200.run_main script.py`: a script that sits next to ANOTHER `networks` package (as the reference's mains
    do) still gets this repo's classes"""
    import subprocess
    import sys
    (tmp_path / "networks").mkdir()
    (tmp_path / "networks" / "__init__.py").write_text("raise ImportError('the shadowed package must not be imported')\n")
    (tmp_path / "utils_local.py").write_text("X = 7\n")
    (tmp_path / "main_x.py").write_text("import sys\nfrom networks import CENet, CENetOrg\nimport utils_local\n"
                                        "print('OK', CENet.__module__, CENetOrg.__module__, utils_local.X, sys.argv[1:])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "cenet_b200.run_main", str(tmp_path / "main_x.py"), "--flag", "1"], cwd=str(tmp_path),
                       env=dict(os.environ, PYTHONPATH=root), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK cenet_b200.networks.cenet cenet_b200.networks.cenet_org 7 ['--flag', '1']" in r.stdout, r.stdout
