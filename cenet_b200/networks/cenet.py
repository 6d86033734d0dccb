"""`networks.CENet` -- drop-in for the reference constructor (src/networks/cenet/net.py:8-64).

The class owns ordinary `nn.Parameter`s / buffers under EXACTLY the reference's state_dict names (801 entries
for the ACDC config), so checkpoints, optimizers, `deepcopy`, `.cuda()`, `.eval()` keep
working.  The parameter tree below is a passive holder: none of the leaf `nn.Module.forward`s is ever called.
`CENet.forward` hands the tensors to `cenet_b200.engine.Engine`, which launches the hand-written sm_100a
kernels of `cenet_b200/csrc` through the C-ABI library.  There is no CPU path: on a CPU tensor, or when the
library is missing, forward raises.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

__all__ = ["CENet"]

_PVT_B2 = dict(embed_dims=(64, 128, 320, 512), heads=(1, 2, 5, 8), mlp_ratios=(8, 8, 4, 4), depths=(3, 4, 6, 3),
               sr_ratios=(8, 4, 2, 1), drop_path_rate=0.1)          # pvtv2.py:400-406
# the PVTv2 variants that share b2's widths (pvtv2.py:392-431): same kernels, other depths / MLP ratios.  b0 (widths 32..256,
# head_dim 32, decoder channels [256,160,64,32]) and the ResNets (encoder.py:34-47) are not built.
PVT_VARIANTS = {
    "pvt_v2_b1": dict(_PVT_B2, depths=(2, 2, 2, 2)),
    "pvt_v2_b2": _PVT_B2,
    "pvt_v2_b3": dict(_PVT_B2, depths=(3, 4, 18, 3)),
    "pvt_v2_b4": dict(_PVT_B2, depths=(3, 8, 27, 3)),
    "pvt_v2_b5": dict(_PVT_B2, depths=(3, 6, 40, 3), mlp_ratios=(4, 4, 4, 4)),
}
_MCA_RATES = {64: (2, 3, 5), 128: (1, 2, 4), 320: (1, 2, 3), 512: (1, 2, 2)}   # decoders.py:64


class _Holder(nn.Module):
    """A named bag of sub-modules / parameters.  Calling it is an error by construction."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the CENet forward runs inside cenet_b200.engine, not per-module")


class _Seq(nn.Sequential):
    """Index-named children (`up_dwc.1.weight`, `dlps.3.2.bias`, ...) with gaps allowed."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k.lstrip("_"), v)


# ---- init helpers (distributions only; pvtv2.py:25-38, blocks.py:97-127, unet.py:112-119) -------------------
def _trunc(w, std=0.02):
    nn.init.trunc_normal_(w, std=std, a=-2.0, b=2.0)


def _init_pvt(m):
    if isinstance(m, nn.Linear):
        _trunc(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def _init_normal(m):
    if isinstance(m, nn.Conv2d):
        nn.init.normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


def _init_unet(m):
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        _trunc(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


# ---- encoder holders ------------------------------------------------------------------------------------------
def _pvt_block(dim, heads, ratio, sr):
    blk = _Holder()
    blk.norm1 = nn.LayerNorm(dim, eps=1e-6)
    attn = _Holder()
    attn.q = nn.Linear(dim, dim, bias=True)
    attn.kv = nn.Linear(dim, 2 * dim, bias=True)
    attn.proj = nn.Linear(dim, dim)
    if sr > 1:
        attn.sr = nn.Conv2d(dim, dim, kernel_size=sr, stride=sr)
        attn.norm = nn.LayerNorm(dim)
    blk.attn = attn
    blk.norm2 = nn.LayerNorm(dim, eps=1e-6)
    mlp = _Holder()
    hid = int(dim * ratio)
    mlp.fc1 = nn.Linear(dim, hid)
    dw = _Holder()
    dw.dwconv = nn.Conv2d(hid, hid, 3, 1, 1, bias=True, groups=hid)
    mlp.dwconv = dw
    mlp.fc2 = nn.Linear(hid, dim)
    blk.mlp = mlp
    return blk


def _pvt_v2(name="pvt_v2_b2", in_chans=3):
    bb = _Holder()
    c = PVT_VARIANTS[name]
    bb.pvt_cfg = c
    prev = in_chans
    for s in range(4):
        pe = _Holder()
        k, st = (7, 4) if s == 0 else (3, 2)
        pe.proj = nn.Conv2d(prev, c["embed_dims"][s], kernel_size=k, stride=st, padding=k // 2)
        pe.norm = nn.LayerNorm(c["embed_dims"][s])
        setattr(bb, f"patch_embed{s+1}", pe)
        prev = c["embed_dims"][s]
    for s in range(4):
        blocks = nn.ModuleList(_pvt_block(c["embed_dims"][s], c["heads"][s], c["mlp_ratios"][s], c["sr_ratios"][s])
                               for _ in range(c["depths"][s]))
        setattr(bb, f"block{s+1}", blocks)
        setattr(bb, f"norm{s+1}", nn.LayerNorm(c["embed_dims"][s], eps=1e-6))
    bb.apply(_init_pvt)
    # stochastic-depth schedule (pvtv2.py:214): train-time only; kept as metadata for the training path
    n = sum(c["depths"])
    bb.drop_path_probs = [c["drop_path_rate"] * i / (n - 1) for i in range(n)]
    return bb


# ---- decoder holders ------------------------------------------------------------------------------------------
def _sep_conv_bn(ch):
    m = _Holder()
    m.depthwise = nn.Conv2d(ch, ch, 3, padding=1, groups=ch, bias=False)
    m.depthwise_bn = nn.BatchNorm2d(ch, eps=1e-5)
    m.pointwise = nn.Conv2d(ch, ch, 1, bias=False)
    m.pointwise_bn = nn.BatchNorm2d(ch, eps=1e-5)
    m.apply(_init_normal)
    return m


def channel_slices(C):
    """cfam.py:178-190, split ratios hard-coded to 5:5:5:1."""
    a, r = int(5 / 16 * C), int(1 / 16 * C)
    return [(0, a), (a, 2 * a), (2 * a, 3 * a), (3 * a, 3 * a + r)]


def _cfa_module(C):
    m = _Holder()
    m.layer_scale_1 = nn.Parameter(1e-6 * torch.ones(1, C, 1, 1))
    m.layer_scale_2 = nn.Parameter(1e-6 * torch.ones(1, C, 1, 1))
    m.norm1 = nn.BatchNorm2d(C, eps=1e-5)
    mca = _Holder()
    mca.gate = nn.Conv2d(C, C, 1)
    val = _Holder()
    sl = channel_slices(C)
    dl = [_sep_conv_bn(b - a) for a, b in sl[:3]]
    ipd = sl[3][1] - sl[3][0]
    dl.append(_Seq(_1=nn.Conv2d(ipd, ipd, 1, bias=False), _2=nn.BatchNorm2d(ipd, eps=1e-5)))
    val.dlps = nn.ModuleList(dl)
    val.PW_conv = nn.Conv2d(C, C, 1)
    mca.value = val
    mca.proj_2 = nn.Conv2d(C, C, 1)
    nl = _Holder()
    nl.w = nn.Parameter(torch.tensor(0.5))
    for n in ("conv_theta", "conv_phi", "conv_g", "conv_out"):
        setattr(nl, n, nn.Conv2d(C, C, 1))
    nl.bn = nn.BatchNorm2d(C, eps=1e-5, momentum=0.1)
    mca.denoising_module = nl
    cc = _Holder()
    cc.fc1 = nn.Conv1d(C, 3 * C, kernel_size=3, groups=C, bias=False)
    cc.fc2 = nn.Conv1d(3 * C, C, kernel_size=1, groups=C, bias=False)
    cc.bn = nn.BatchNorm1d(C)
    mca.ccu = cc
    m.mca = mca
    m.norm2 = nn.BatchNorm2d(C, eps=1e-5)
    mlp = _Holder()
    mlp.fc1 = nn.Conv2d(C, 4 * C, 1)
    mlp.dwconv = nn.Conv2d(4 * C, 4 * C, 3, padding=1, groups=4 * C, bias=True)
    mlp.fc2 = nn.Conv2d(4 * C, C, 1)
    s = _Holder()
    s.pwc = nn.Conv2d(3, 1, 1, bias=False)
    s.dwc = nn.Conv2d(3, 1, 3, padding=1, bias=False)
    s.bn = nn.BatchNorm2d(1)
    mlp.srm = s
    m.mlp = mlp
    return m


def _eucb(cin, cout):
    m = _Holder()
    m.up_dwc = _Seq(_1=nn.Conv2d(cin, cin, 3, padding=1, groups=cin, bias=False), _2=nn.BatchNorm2d(cin))
    m.pwc = _Seq(_0=nn.Conv2d(cin, cout, 1, bias=True))
    m.apply(_init_normal)
    m.kind = "eucb"
    return m


def _up_conv(cin, cout, ks=3):
    m = _Holder()
    m.up = _Seq(_1=nn.Conv2d(cin, cout, ks, padding=ks // 2, bias=False), _2=nn.BatchNorm2d(cout))
    m.apply(_init_normal)
    m.kind = "upcn"
    return m


def _make_up(kind, cin, cout, ks=3):
    if kind not in ("uprb", "eucb", "upcn", "uptc"):
        raise AssertionError(f"Invalid up_block: {kind}")                  # decoders.py:47 / out.py:30
    if kind == "eucb":
        return _eucb(cin, cout)
    if kind == "upcn":
        return _up_conv(cin, cout, ks)
    if kind == "uprb":                                                      # blocks.py:188-204
        m = _Holder()
        m.up = _Seq(_1=_res_block(cin, cout, ks))
        m.kind = "uprb"
        return m
    # uptc (blocks.py:223-243): monai Convolution(conv_only=True, is_transposed=True) == Sequential(conv=ConvTranspose2d) with
    # padding (k - s + 1) // 2 and output_padding 2p + s - k (unet.py:16-48); the decoder calls it with stride 2
    m = _Holder()
    pad = (ks - 2 + 1) // 2
    m.up = _Seq(conv=nn.ConvTranspose2d(cin, cout, ks, stride=2, padding=pad, output_padding=2 * pad + 2 - ks, bias=False))
    m.apply(_init_normal)
    m.kind = "uptc"
    return m


def _dse_block(dim, scale_factors, heads, depth, mode="cat"):
    m = _Holder()
    E = 2 * dim if mode == "cat" else dim                                  # dseb.py:96
    b = _Holder()
    b.w = nn.Parameter(torch.randn(1, E, 1, 1) + 0.5)                      # dseb.py:35
    m.boundary = b
    d = _Holder()
    hd = E // heads // 2
    for n in ("lambda_q1", "lambda_k1", "lambda_q2", "lambda_k2"):         # multihead_diffattn.py:63-66
        setattr(d, n, nn.Parameter(torch.zeros(hd).normal_(0, 0.1)))
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        setattr(d, n, nn.Linear(E, E, bias=False))
    m.diffattn = d
    m.mixer = nn.Conv2d(E, dim, 1, bias=False)
    m.meta = dict(heads=heads, depth=depth, scale_factors=list(scale_factors))
    return m


def _decoder(channels, scale_factors, heads, up_block, skip_mode="cat"):
    d = _Holder()
    d.dec4 = _cfa_module(channels[0])
    d.up3 = _make_up(up_block, channels[0], channels[1])
    d.skip_enhancer3 = _dse_block(channels[1], scale_factors, heads[0], 4, skip_mode)
    d.dec3 = _cfa_module(channels[1])
    d.up2 = _make_up(up_block, channels[1], channels[2])
    d.skip_enhancer2 = _dse_block(channels[2], scale_factors, heads[1], 3, skip_mode)
    d.dec2 = _cfa_module(channels[2])
    d.up1 = _make_up(up_block, channels[2], channels[3])
    d.skip_enhancer1 = _dse_block(channels[3], scale_factors, heads[2], 2, skip_mode)
    d.dec1 = _cfa_module(channels[3])
    return d


def _conv_only(cin, cout, k, bias=False):
    return _Seq(conv=nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias))


def _res_block(cin, cout, k):
    m = _Holder()
    m.conv1 = _conv_only(cin, cout, k)
    m.conv2 = _conv_only(cout, cout, k)
    m.norm1 = nn.BatchNorm2d(cout)
    m.norm2 = nn.BatchNorm2d(cout)
    if cin != cout:
        m.conv3 = _conv_only(cin, cout, 1)
        m.norm3 = nn.BatchNorm2d(cout)
    m.apply(_init_unet)
    return m


def _out_head(dec_ch, x_ch, ncls, merge_mode, up_block, up_ks):
    if merge_mode not in ("cat", "add"):
        raise AssertionError(f"Invalid merge_mode: {merge_mode}")
    o = _Holder()
    om = dec_ch // 2
    mix = om if merge_mode == "add" else 2 * om                            # out.py:47
    o.w = nn.Parameter(torch.randn(1, om, 1, 1) + 0.75)                    # out.py:39
    ob = _Holder()
    ob.conv = _conv_only(mix, ncls, 1, bias=True)
    ob.apply(_init_unet)
    o.out = _Seq(_0=_res_block(mix, mix, 3), _1=ob)
    o.up = _make_up(up_block, dec_ch, om, up_ks)
    o.rb = _Seq(_0=_res_block(x_ch, om, 5))
    return o


class _TrainForward(torch.autograd.Function):
    """autograd boundary of the hand-written training path: `net(x)` in train() mode returns logits whose backward runs
    the recorded backward kernels of cenet_b200.train.TrainEngine and hands the parameter gradients to autograd, so the
    reference loop (`loss = criterion(net(x), y); loss.backward(); optimizer.step()`, main_acdc.py:250-262) runs unchanged."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.train_engine(x.device)
        ctx.eng = eng
        ctx.names = [n for n, _ in module.named_parameters()]
        ctx.needs = [p.requires_grad for p in params]
        return eng.forward_logits(x).clone()

    @staticmethod
    def backward(ctx, dlogits):
        eng = ctx.eng
        eng.backward_from(dlogits.contiguous())
        flat = eng.gflat.clone()                                      # ONE copy; the per-parameter gradients are views of it
        grads = tuple(flat[eng.param_offsets[n]:eng.param_offsets[n] + eng.GP[n].numel()].view(eng.GP[n].shape) if need else None
                      for n, need in zip(ctx.names, ctx.needs))
        return (None, None) + grads


class _EvalForward(torch.autograd.Function):
    """eval()-mode forward under grad mode (ACDC `val()` main_acdc.py:226, utils_skin.py:104,143 and the FLOP-counter warm-ups
    utils.py:184-185 run the network that way and never call backward): the logits come from the inference launch plan and carry
    a graph node like the reference's output does (`requires_grad`, `.detach()`, `.item()` behave the same).  A backward pass through
    the eval-mode network recomputes the forward with the training launch plan in frozen-statistics mode (see `backward`)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        ctx.module, ctx.x = module, x.detach()
        ctx.names = [n for n, _ in module.named_parameters()]
        ctx.needs = [p.requires_grad for p in params]
        return module._engine(x).forward(x)

    @staticmethod
    def backward(ctx, dlogits):
        # Rare path (the reference scripts never call it): recompute the forward with the training launch plan in FROZEN-statistics
        # mode -- BatchNorm / BatchNorm1d(CCU) / BatchNorm2d(1)(SRM) use their running statistics and update nothing, DropPath is
        # the identity, i.e. exactly eval() semantics -- then run its recorded backward.  Eager, no CUDA graphs.
        eng = ctx.module.train_engine(ctx.x.device)
        saved = (eng.frozen_stats, eng.use_graph)
        eng.frozen_stats, eng.use_graph = True, False
        try:
            eng.forward_logits(ctx.x)
            eng.backward_from(dlogits.contiguous())
        finally:
            eng.frozen_stats, eng.use_graph = saved
        flat = eng.gflat.clone()
        grads = tuple(flat[eng.param_offsets[n]:eng.param_offsets[n] + eng.GP[n].numel()].view(eng.GP[n].shape) if need else None
                      for n, need in zip(ctx.names, ctx.needs))
        return (None, None) + grads


class CENet(nn.Module):
    """Same constructor as the reference `networks.CENet` (net.py:9-22)."""

    def __init__(self, input_channels=1, num_classes=1, scale_factors=[0.8, 0.4], diffatt_num_heads=[2, 2, 2],
                 encoder="pvt_v2_b2", enc_pretrain=False, freeze_bb=False, skip_mode="cat", dec_up_block="eucb",
                 out_merge_mode="cat", out_up_block="eucb", out_up_ks=3, writer=None, base_ptdir="."):
        super().__init__()
        self.writer = writer
        num_classes = int(num_classes)
        if encoder == "pvt_v2_b0" or "resnet" in str(encoder):
            raise NotImplementedError(f"encoder '{encoder}': the accelerated path covers pvt_v2_b1 .. b5 (the variants with the "
                                      "widths 64/128/320/512, encoder.py:14-33); see DESIGN.md section 7")
        path = f"{base_ptdir}/pvt/{encoder}.pth"                            # encoder.py:15-33
        if encoder not in PVT_VARIANTS:                                     # encoder.py:48-52 silent fallback
            print("Encoder not implemented! Continuing with default encoder pvt_v2_b2.")
            encoder = "pvt_v2_b2"
            path = f"{base_ptdir}/pretrained_pth/pvt/pvt_v2_b2.pth"
        skip_mode = skip_mode.lower()
        if skip_mode not in ("cat", "add"):                                 # (the reference treats anything but 'add' as cat)
            skip_mode = "cat"
        channels = [512, 320, 128, 64]
        self.backbone = _pvt_v2(encoder, 3)
        if enc_pretrain and base_ptdir:                                     # encoder.py:73-84
            print(f"Loading pretrained weights from {path}")
            saved = torch.load(path)
            own = self.backbone.state_dict()
            own.update({k: v for k, v in saved.items() if k in own})
            self.backbone.load_state_dict(own)
            if freeze_bb:
                for p in self.backbone.parameters():
                    p.requires_grad = False
        else:
            print("No pretrained weights loaded! ...")
        self.decoder = _decoder(channels, scale_factors, diffatt_num_heads, dec_up_block, skip_mode)
        self.out = _out_head(channels[-1], input_channels, num_classes, out_merge_mode, out_up_block, out_up_ks)
        self.cfg = dict(input_channels=input_channels, num_classes=num_classes, scale_factors=list(scale_factors),
                        diffatt_num_heads=list(diffatt_num_heads), dec_up_block=dec_up_block,
                        out_up_block=out_up_block, encoder=encoder, skip_mode=skip_mode, out_merge_mode=out_merge_mode)
        self._engines = {}

    # engines hold device workspaces + packed weights; they are rebuilt lazily and never copied / pickled
    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        return d

    def _replicate_for_data_parallel(self):
        # nn.DataParallel (main_acdc.py:177-178, a dead branch in the scripts: `--n_gpu` is never passed) shallow-copies the
        # module per device and per iteration; the launch plans cache packed weights / CUDA graphs per module instance and
        # the replicas' tensors are not leaf Parameters, so results would silently be computed with stale weights.
        raise RuntimeError("cenet_b200.CENet does not support nn.DataParallel replicas: run one process per GPU "
                           "(torchrun) and attach cenet_b200.replicas.GradSync(net.train_engine()) -- see INTEGRATION.md")

    def _engine(self, x, precision=None):
        from ..engine import Engine
        key = (x.device, precision or Engine.default_precision())
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = Engine(self, x.device, key[1])
        return eng

    def train_engine(self, device=None, precision=None):
        """The training launch plan of this module on `device` (cenet_b200.train.TrainEngine); created on first use.
        It moves the parameters into one flat fp32 buffer (they stay ordinary nn.Parameters, now views)."""
        from ..engine import Engine
        from ..train import TrainEngine
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        key = ("train", dev, precision or Engine.default_precision())
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = TrainEngine(self, dev, key[2])
        return eng

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("cenet_b200.CENet has no CPU path: move the module and the input to a B200 "
                               "(`.cuda()`); the CPU oracle lives in oracle/ and is test-only")
        if self.training:
            if torch.is_grad_enabled():
                params = [p for p in self.parameters()]
                return _TrainForward.apply(self, x, *params)
            return self.train_engine(x.device).forward_logits(x).clone()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _EvalForward.apply(self, x, *self.parameters())
        return self._engine(x).forward(x)

    @torch.no_grad()
    def predict(self, x, out=None):
        """logits -> `argmax(softmax(.,1),1)` fused on the device (metrics_eval.py:52); int64 [B,H,W].
        out: optional preallocated int64 [B,H,W] CUDA tensor (a serving loop that overlaps the D2H copy of the previous
        batch with this call rotates two of them)."""
        return self._engine(x).forward(x, labels=True, out=out)
