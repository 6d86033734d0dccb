"""Generates tests/golden/*.pt by running the REAL reference (/root/reference/src/networks, imported through
oracle/ref_shim.py) in the build container.  Re-run with:  python tests/golden/make_golden.py

Each fixture stores what is needed to re-create the inputs deterministically elsewhere (config name, seeds; the
weights come from `cenet_b200.networks.CENet` built under torch.manual_seed(seed) and passed through
oracle.fixtures.perturb_state -- loaded into the reference with strict=True, which also pins the 801-key
state_dict contract) plus the reference's outputs: strided logits, label histogram and per-module tap statistics.
Small module-level fixtures (DiffAttn, FEA, Nonlocal, CCU, SRM, ...) store full tensors.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures, ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def build_pair(nets, name, seed=1234):
    from cenet_b200.networks import CENet
    kw = fixtures.CONFIGS[name]
    torch.manual_seed(seed)
    mine = CENet(**kw)
    sd = fixtures.perturb_state(mine.state_dict(), seed)
    ref = nets.CENet(**kw).eval()
    ref.load_state_dict(sd, strict=True)
    return ref, sd, kw


def model_fixture(nets, name, batch):
    ref, sd, kw = build_pair(nets, name)
    x = fixtures.synth_input(name, batch)
    taps = {}
    hooks = []
    for mod_name in ["decoder.dec4", "decoder.dec3", "decoder.dec2", "decoder.dec1", "decoder.up3", "decoder.up2",
                     "decoder.up1", "decoder.skip_enhancer3", "decoder.skip_enhancer2", "decoder.skip_enhancer1"]:
        m = ref.get_submodule(mod_name)
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=mod_name: taps.__setitem__(n, o.detach())))
    with torch.no_grad():
        feats = ref.backbone(torch.cat([x, x, x], 1) if x.shape[1] == 1 else x)
        y = ref(x)
    for h in hooks:
        h.remove()
    for i, f in enumerate(feats):
        taps[f"backbone.stage{i+1}"] = f
    lab = torch.argmax(torch.softmax(y, 1), 1)
    out = dict(config=name, batch=batch, seed=1234, input_seed=0,
               logits_strided=y[:, :, ::8, ::8].clone(), logits_mean=y.mean().item(), logits_std=y.std().item(),
               logits_abs_sum=y.abs().sum().item(),
               label_hist=torch.bincount(lab.flatten(), minlength=kw["num_classes"]),
               labels_strided=lab[:, ::4, ::4].clone(),
               taps={k: dict(mean=v.mean().item(), std=v.std().item(), sample=v.flatten()[:: max(1, v.numel() // 512)][:512].clone())
                     for k, v in taps.items()})
    torch.save(out, os.path.join(HERE, f"model_{name}_b{batch}.pt"))
    print(name, batch, "logits std", out["logits_std"], "hist", out["label_hist"].tolist())


def module_fixtures(nets):
    """Small full-tensor fixtures of the reference's own sub-modules (seeded default init + perturbed BN stats)."""
    import networks.cenet.modules.cfam as cfam
    import networks.cenet.modules.dseb as dseb
    import networks.cenet.modules.multihead_diffattn as mda
    import networks.cenet.modules.nlb as nlb
    import networks.cenet.modules.blocks as blocks
    import networks.cenet.pvtv2 as pvt
    import networks.cenet.out as outm
    out = {}

    def run(key, mod, *inputs):
        mod.eval()
        sd = fixtures.perturb_state(mod.state_dict(), 7)
        mod.load_state_dict(sd)
        with torch.no_grad():
            y = mod(*inputs)
        out[key] = dict(state=sd, inputs=[i.clone() if torch.is_tensor(i) else i for i in inputs], output=y.clone())

    g = torch.Generator().manual_seed(11)
    torch.manual_seed(11)
    run("diffattn_e64_h2_n80", mda.MultiheadDiffAttn(64, depth=2, num_heads=2), torch.randn(2, 80, 64, generator=g))
    run("diffattn_e32_h2_n64", mda.MultiheadDiffAttn(32, depth=3, num_heads=2), torch.randn(2, 64, 32, generator=g) * 2)
    run("fea_2scales", dseb.FEA(16, [0.8, 0.4]), torch.randn(2, 16, 14, 14, generator=g))
    run("fea_3scales", dseb.FEA(8, [1.0, 0.75, 0.5]), torch.randn(1, 8, 28, 28, generator=g))
    run("dseb_c16", dseb.DSEBlock(16, [0.8, 0.4], 2, 14, mode="cat", depth=3),
        torch.randn(2, 16, 14, 14, generator=g), torch.randn(2, 16, 14, 14, generator=g))
    run("nonlocal_c64", nlb.Nonlocal(64), torch.randn(2, 64, 14, 14, generator=g))
    run("ccu_c64_b2", cfam.CCU(64), torch.randn(2, 64, 14, 14, generator=g))
    run("ccu_c64_b1", cfam.CCU(64), torch.randn(1, 64, 14, 14, generator=g))
    run("srm", cfam.SRM(), torch.randn(2, 64, 14, 14, generator=g))
    run("modw_c64", cfam.MultiOrderDWConv(64, rates=[2, 3, 5]), torch.randn(2, 64, 28, 28, generator=g))
    run("cfam_c64", cfam.CFAModule(64, ffn_ratio=4, drop_rate=0, drop_path_rate=0, act_type="GELU", norm_type="BN",
                                   init_value=1e-6, attn_channel_split=[1, 3, 4], attn_act_type="SiLU",
                                   mca_rates=[2, 3, 5]), torch.randn(2, 64, 14, 14, generator=g))
    run("eucb", blocks.EUCB(64, 32, kernel_size=3, stride=1, activation="leakyrelu"), torch.randn(2, 64, 7, 7, generator=g))
    run("upconv", blocks.UpConv(64, 32, kernel_size=3, stride=1, activation="leakyrelu"), torch.randn(2, 64, 7, 7, generator=g))
    run("pvt_attn_sr2", pvt.Attention(128, num_heads=2, qkv_bias=True, sr_ratio=2), torch.randn(2, 196, 128, generator=g), 14, 14)
    run("pvt_block_sr1", pvt.Block(64, 1, mlp_ratio=4, qkv_bias=True, sr_ratio=1), torch.randn(2, 49, 64, generator=g), 7, 7)
    run("outhead", outm.OutHead(dec_in_channels=64, x_in_channels=1, out_channels=4, up_block="upcn"),
        torch.randn(1, 64, 16, 16, generator=g), torch.randn(1, 1, 64, 64, generator=g))
    torch.save(out, os.path.join(HERE, "modules.pt"))
    print("modules:", list(out))


def loss_fixture():
    import types
    import importlib.util
    # utils/utils.py pulls thop/ptflops/fvcore/matplotlib (absent here); core.py only needs `flatten` from it, which
    # the dice/ce path never calls -> give core.py a stub parent package and load the file itself unmodified
    pkg = types.ModuleType("refutils")
    pkg.__path__ = []
    stub = types.ModuleType("refutils.utils")
    stub.flatten = None
    sys.modules["refutils"], sys.modules["refutils.utils"] = pkg, stub
    spec = importlib.util.spec_from_file_location("refutils.core", "/root/reference/src/utils/core.py")
    core = importlib.util.module_from_spec(spec)
    sys.modules["refutils.core"] = core
    spec.loader.exec_module(core)
    g = torch.Generator().manual_seed(5)
    out = {}
    for ncls, B, S in ((4, 2, 32), (9, 3, 24), (2, 1, 40)):
        logits = torch.randn(B, ncls, S, S, generator=g) * 2
        logits.requires_grad_(True)
        labels = torch.randint(0, ncls, (B, S, S), generator=g).float()
        args = types.SimpleNamespace(loss_type="dice,ce", loss_weights="0.5,0.5")
        crit = core.Criterion(ncls, args)
        loss = crit(logits, labels)
        loss.backward()
        out[f"c{ncls}"] = dict(logits=logits.detach().clone(), labels=labels.long(), loss=loss.detach().clone(),
                               grad=logits.grad.clone())
    torch.save(out, os.path.join(HERE, "loss.pt"))
    print("loss:", {k: v["loss"].item() for k, v in out.items()})
    # BoundaryDoULoss (core.py:83-131) hard-codes `.cuda()` on its scratch tensors; on this GPU-less container the call is
    # patched to the identity for the duration of the run (arithmetic untouched).  Labels: smooth blobs (nearest-upsampled
    # coarse maps) so that interior and boundary pixels both exist, and pure noise (every pixel a boundary pixel).
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        outb = {}
        for ncls, B, S, blob in ((4, 2, 32, 4), (9, 3, 24, 3), (2, 1, 40, 8), (4, 2, 16, 1)):
            coarse = torch.randint(0, ncls, (B, 1, S // blob, S // blob), generator=g).float()
            labels = torch.nn.functional.interpolate(coarse, scale_factor=blob, mode="nearest")[:, 0]
            for lt, lw in (("boundary", "1.0"), ("dice,ce,boundary", "0.3,0.3,0.4")):
                logits = (torch.randn(B, ncls, S, S, generator=g) * 2).requires_grad_(True)
                crit = core.Criterion(ncls, types.SimpleNamespace(loss_type=lt, loss_weights=lw))
                loss = crit(logits, labels)
                loss.backward()
                outb[f"c{ncls}_s{S}_{lt}"] = dict(logits=logits.detach().clone(), labels=labels.long(), loss=loss.detach().clone(),
                                                 grad=logits.grad.clone(), loss_type=lt, loss_weights=lw)
    finally:
        torch.Tensor.cuda = real_cuda
    torch.save(outb, os.path.join(HERE, "loss_boundary.pt"))
    print("boundary loss:", {k: round(v["loss"].item(), 6) for k, v in outb.items()})


TRAIN_GRAD_KEYS = [       # >= 1 tensor from every module family; full gradient norms of all 630 parameters are stored too
    "backbone.patch_embed1.proj.weight", "backbone.patch_embed3.norm.weight", "backbone.block1.0.attn.q.weight",
    "backbone.block1.0.attn.sr.weight", "backbone.block1.1.attn.norm.bias", "backbone.block2.1.mlp.dwconv.dwconv.weight",
    "backbone.block2.3.attn.kv.weight", "backbone.block3.2.mlp.fc1.weight", "backbone.block3.5.attn.proj.bias",
    "backbone.block4.0.attn.kv.bias", "backbone.block4.2.mlp.fc2.weight", "backbone.norm4.weight",
    "decoder.dec4.layer_scale_1", "decoder.dec4.mca.ccu.fc1.weight", "decoder.dec4.mca.ccu.bn.weight",
    "decoder.dec3.mca.gate.weight", "decoder.dec3.mca.value.dlps.1.depthwise.weight", "decoder.dec3.mca.value.dlps.1.pointwise_bn.bias",
    "decoder.dec2.mca.value.dlps.3.1.weight", "decoder.dec2.mca.value.PW_conv.weight", "decoder.dec2.mca.proj_2.bias",
    "decoder.dec1.mca.denoising_module.w", "decoder.dec1.mca.denoising_module.conv_theta.weight",
    "decoder.dec1.mca.denoising_module.bn.weight", "decoder.dec1.norm2.weight", "decoder.dec1.mlp.dwconv.weight",
    "decoder.dec1.mlp.srm.pwc.weight", "decoder.dec1.mlp.srm.dwc.weight", "decoder.dec1.mlp.srm.bn.bias", "decoder.dec1.layer_scale_2",
    "decoder.up3.up_dwc.1.weight", "decoder.up2.up_dwc.2.weight", "decoder.up1.pwc.0.bias",
    "decoder.skip_enhancer1.boundary.w", "decoder.skip_enhancer1.diffattn.lambda_q1", "decoder.skip_enhancer1.diffattn.lambda_k2",
    "decoder.skip_enhancer1.diffattn.v_proj.weight", "decoder.skip_enhancer2.diffattn.q_proj.weight",
    "decoder.skip_enhancer3.diffattn.out_proj.weight", "decoder.skip_enhancer2.mixer.weight",
    "out.w", "out.rb.0.conv1.conv.weight", "out.rb.0.conv2.conv.weight", "out.rb.0.conv3.conv.weight", "out.rb.0.norm3.weight",
    "out.up.up.1.weight", "out.up.up.2.bias", "out.out.0.conv1.conv.weight", "out.out.0.norm2.weight",
    "out.out.1.conv.conv.weight", "out.out.1.conv.conv.bias",
]


def _sample(t, n=2048):
    f = t.detach().flatten()
    return f[:: max(1, f.numel() // n)][:n].clone()


def train_fixture(nets, name, batch, size):
    """The reference in train() mode (batch-statistics BatchNorm with running-stat updates, the CCU `B > 1` guard, DropPath
    with drop_prob forced to 0), one forward + `Criterion('dice,ce', 0.5/0.5)` (the reference's own utils/core.py class) +
    backward: loss, strided logits, gradients (norm of all 630, samples of TRAIN_GRAD_KEYS) and every BatchNorm buffer
    after the step.  Pins oracle.cenet_forward(training=True) + autograd, which every gradient test compares against."""
    import importlib.util
    import types
    ref, sd, kw = build_pair(nets, name)
    ref.train()
    n_dp = 0
    for m in ref.modules():
        if type(m).__name__ == "DropPath":
            m.drop_prob = 0.0
            n_dp += 1
    if size != 224:                                            # dseb.py:117 recovers H from a constructor constant
        for lvl, div in ((3, 16), (2, 8), (1, 4)):
            getattr(ref.decoder, f"skip_enhancer{lvl}").input_size = size // div
    pkg = types.ModuleType("refutils")
    pkg.__path__ = []
    stub = types.ModuleType("refutils.utils")
    stub.flatten = None
    sys.modules["refutils"], sys.modules["refutils.utils"] = pkg, stub
    spec = importlib.util.spec_from_file_location("refutils.core", "/root/reference/src/utils/core.py")
    core = importlib.util.module_from_spec(spec)
    sys.modules["refutils.core"] = core
    spec.loader.exec_module(core)
    x = fixtures.synth_input(name, batch, size=size)
    labels = torch.randint(0, kw["num_classes"], (batch, size, size), generator=torch.Generator().manual_seed(5))
    crit = core.Criterion(kw["num_classes"], types.SimpleNamespace(loss_type="dice,ce", loss_weights="0.5,0.5"))
    y = ref(x)
    loss = crit(y, labels.float())
    loss.backward()
    grads = {n: p.grad for n, p in ref.named_parameters()}
    after = ref.state_dict()
    out = dict(config=name, batch=batch, size=size, seed=1234, input_seed=0, label_seed=5, drop_path_modules=n_dp,
               loss=loss.detach().clone(), logits_strided=y.detach()[:, :, ::4, ::4].clone(),
               logits_std=y.std().item(),
               grad_norms={n: (g.norm().item() if g is not None else None) for n, g in grads.items()},
               grad_samples={n: _sample(grads[n]) for n in TRAIN_GRAD_KEYS if grads.get(n) is not None},
               buffers_after={k: v.clone() for k, v in after.items() if "running_" in k or k.endswith("num_batches_tracked")})
    torch.save(out, os.path.join(HERE, f"train_{name}_b{batch}_s{size}.pt"))
    none = [n for n, g in grads.items() if g is None]
    print("train", name, batch, size, "loss", loss.item(), "params without grad:", none)


ORG_KW = dict(num_classes=9, input_channels=1, scale_factors=[0.8, 0.4], encoder="pvt_v2_b2", pretrain=False,
              num_heads=[16, 8, 8])           # scripts/synapse.sh TEST_ORG -> main_synapse.py:129-137


def org_fixture(nets, batch):
    """CENetOrg (the reference's cenet_org.net.Net) in eval mode: strided logits, labels and per-module taps.  The weights come
    from cenet_b200.networks.CENetOrg and are loaded into the reference with strict=True (pins the 822-key contract)."""
    from cenet_b200.networks import CENetOrg
    torch.manual_seed(1234)
    mine = CENetOrg(**ORG_KW)
    sd = fixtures.perturb_state(mine.state_dict(), 1234)
    sd["out.conv.conv.weight"] = sd["out.conv.conv.weight"] * 20.0      # trained-like logit margins, as for CENet
    ref = nets.CENetOrg(**ORG_KW).eval()
    ref.load_state_dict(sd, strict=True)
    x = fixtures.synth_input("synapse", batch)
    taps, hooks = {}, []
    for mod_name in ["decoder.dec4", "decoder.dec3", "decoder.dec2", "decoder.dec1", "decoder.eucb3", "decoder.eucb2",
                     "decoder.eucb1", "decoder.skip_enhancer3", "decoder.skip_enhancer2", "decoder.skip_enhancer1"]:
        m = ref.get_submodule(mod_name)
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=mod_name: taps.__setitem__(n, o.detach())))
    with torch.no_grad():
        feats = ref.backbone(ref.conv(x))
        y = ref(x)
    for h in hooks:
        h.remove()
    for i, f in enumerate(feats):
        taps[f"backbone.stage{i+1}"] = f
    lab = torch.argmax(torch.softmax(y, 1), 1)
    out = dict(kw=ORG_KW, batch=batch, seed=1234, input_seed=0, n_keys=len(sd),
               logits_strided=y[:, :, ::4, ::4].clone(), logits_std=y.std().item(),
               label_hist=torch.bincount(lab.flatten(), minlength=9), labels_strided=lab[:, ::2, ::2].clone(),
               taps={k: dict(mean=v.mean().item(), std=v.std().item(), norm=v.norm().item(),
                             sample=v.flatten()[:: max(1, v.numel() // 4096)][:4096].clone()) for k, v in taps.items()})
    torch.save(out, os.path.join(HERE, f"model_org_synapse_b{batch}.pt"))
    print("org", batch, "keys", len(sd), "logits std", out["logits_std"], "hist", out["label_hist"].tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "org":
        nets = ref_shim.import_reference()
        org_fixture(nets, 2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":         # only the train-mode fixtures
        nets = ref_shim.import_reference()
        train_fixture(nets, "acdc", 2, 224)
        train_fixture(nets, "synapse", 2, 96)
        train_fixture(nets, "acdc", 1, 64)                   # B = 1 in train mode: CCU skips BatchNorm1d (cfam.py:260)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "loss":          # only the loss fixtures (no model import needed)
        loss_fixture()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":      # other PVTv2 encoders
        nets = ref_shim.import_reference()
        model_fixture(nets, "acdc_b1", 1)
        model_fixture(nets, "acdc_b5", 1)
        model_fixture(nets, "acdc_add", 1)
        model_fixture(nets, "synapse_uprb", 1)
        model_fixture(nets, "acdc_uptc", 1)
        sys.exit(0)
    nets = ref_shim.import_reference()
    module_fixtures(nets)
    loss_fixture()
    model_fixture(nets, "acdc", 1)
    model_fixture(nets, "synapse", 2)
    model_fixture(nets, "skin", 1)
    train_fixture(nets, "acdc", 2, 224)
    train_fixture(nets, "synapse", 2, 96)
    train_fixture(nets, "acdc", 1, 64)
