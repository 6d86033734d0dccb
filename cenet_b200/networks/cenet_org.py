"""`networks.CENetOrg` -- drop-in for the reference's "original paper" variant (src/networks/cenet_org/net.py:16-129).

main_synapse.py:15 imports the symbol and `scripts/synapse.sh TEST_ORG` evaluates a published checkpoint with it:
`CENetOrg(input_channels, num_classes, scale_factors=[0.8, 0.4], encoder='pvt_v2_b2', pretrain=True, num_heads=[16, 8, 8])`.
Same PVTv2-b2 encoder as CENet; the decoder differs in the wiring, not in the primitives (cenet_org/decoders.py:112-197):
  * gray -> 3 channels through Conv1x1 + BatchNorm + ReLU (net.py:23-28) instead of `cat([x, x, x])`
  * SkipEnhancer: DoGEdge `y + w |up(down_s0 y) - up(down_s1 y)|` BEFORE the differential attention (depth 1), the gate is
    applied in token space (`diffattn(tok) * tok`, which is the same flat multiplication), `z = y + gated`, 1x1 `proj` WITH bias
  * CFAMBlock = CFAModule with fixed dilation rates 6/12/18, ReLU in the image-pooling branch; EUCB with ReLU
  * head: `enc` = UnetResBlock(Cin -> 32, k=3) + MaxPool2 (no learnable scale), `up` = bilinear x2 + UnetResBlock(64 -> 32, k=3),
    `rb` = UnetResBlock(64 -> 64, k=3), 1x1 `out`, bilinear x2
The class owns ordinary nn.Parameters under the reference's 822 state_dict names (so the published checkpoint layout loads
with strict=True); `forward` hands them to `cenet_b200.engine_org.EngineOrg`.  Inference only: the reference's scripts never
train this variant (`--model_version cenet_org` is only used with `--eval`), train() mode raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .cenet import (_Holder, _Seq, _conv_only, _eucb, _init_normal, _init_unet, _pvt_v2, _res_block, _sep_conv_bn,
                    channel_slices)

__all__ = ["CENetOrg"]


def _cfam_block(C):
    """cenet_org/modules/cfam.py:336-391 (CFAMBlock): CFAModule with `attn` / `crm` attribute names"""
    m = _Holder()
    m.layer_scale_1 = nn.Parameter(1e-6 * torch.ones(1, C, 1, 1))
    m.layer_scale_2 = nn.Parameter(1e-6 * torch.ones(1, C, 1, 1))
    m.norm1 = nn.BatchNorm2d(C, eps=1e-5)
    at = _Holder()
    at.gate = nn.Conv2d(C, C, 1)
    val = _Holder()
    sl = channel_slices(C)
    dl = [_sep_conv_bn(b - a) for a, b in sl[:3]]
    ipd = sl[3][1] - sl[3][0]
    dl.append(_Seq(_1=nn.Conv2d(ipd, ipd, 1, bias=False), _2=nn.BatchNorm2d(ipd, eps=1e-5)))
    val.dlps = nn.ModuleList(dl)
    val.PW_conv = nn.Conv2d(C, C, 1)
    at.value = val
    at.proj_2 = nn.Conv2d(C, C, 1)
    nl = _Holder()
    nl.w = nn.Parameter(torch.tensor(0.5))
    for n in ("conv_theta", "conv_phi", "conv_g", "conv_out"):
        setattr(nl, n, nn.Conv2d(C, C, 1))
    nl.bn = nn.BatchNorm2d(C, eps=1e-5, momentum=0.1)
    at.denoising_module = nl
    cr = _Holder()
    cr.fc1 = nn.Conv1d(C, 3 * C, kernel_size=3, groups=C, bias=False)
    cr.fc2 = nn.Conv1d(3 * C, C, kernel_size=1, groups=C, bias=False)
    cr.bn = nn.BatchNorm1d(C)
    at.crm = cr
    m.attn = at
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


def _skip_enhancer(dim, heads):
    """cenet_org/decoders.py:128-144, mode='cat'"""
    m = _Holder()
    E = 2 * dim
    b = _Holder()
    b.w = nn.Parameter(torch.ones(1, E, 1, 1) * 0.5)
    m.boundary = b
    d = _Holder()
    hd = E // heads // 2
    for n in ("lambda_q1", "lambda_k1", "lambda_q2", "lambda_k2"):
        setattr(d, n, nn.Parameter(torch.zeros(hd).normal_(0, 0.1)))
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        setattr(d, n, nn.Linear(E, E, bias=False))
    m.diffattn = d
    m.proj = nn.Conv2d(E, dim, 1)
    m.meta = dict(heads=heads, depth=1)
    return m


class CENetOrg(nn.Module):
    """Same constructor as the reference `networks.CENetOrg` (= cenet_org.net.Net, net.py:17-19)."""

    def __init__(self, num_classes=1, input_channels=1, scale_factors=[0.6, 0.3], num_heads=[2, 2, 2],
                 encoder="pvt_v2_b2", pretrain=False, skip_mode="cat", base_ptdir="."):
        super().__init__()
        num_classes = int(num_classes)
        if encoder in ("pvt_v2_b0", "pvt_v2_b1", "pvt_v2_b3", "pvt_v2_b4", "pvt_v2_b5") or "resnet" in str(encoder):
            raise NotImplementedError(f"CENetOrg encoder '{encoder}': only pvt_v2_b2 (the published configuration) is built")
        if encoder != "pvt_v2_b2":                                          # net.py:69-73 silent fallback
            print("Encoder not implemented! Continuing with default encoder pvt_v2_b2.")
        if str(skip_mode).lower() != "cat":
            raise NotImplementedError("CENetOrg skip_mode='add': only 'cat' (the published configuration) is built")
        if len(scale_factors) != 2:
            raise ValueError("DoGEdge takes exactly two scale factors (cenet_org/decoders.py:118-119)")
        if input_channels == 1:
            self.conv = nn.Sequential(nn.Conv2d(1, 3, kernel_size=1), nn.BatchNorm2d(3), nn.ReLU(inplace=True))
        else:
            self.conv = nn.Identity()
        self.backbone = _pvt_v2("pvt_v2_b2", 3)
        if pretrain:                                                        # net.py:75-84 (failure is only printed)
            path = f"{base_ptdir}/pretrained_pth/pvt/pvt_v2_b2.pth"
            try:
                saved = torch.load(path)
                own = self.backbone.state_dict()
                own.update({k: v for k, v in saved.items() if k in own})
                self.backbone.load_state_dict(own)
                print(f"Loaded pretrained weights from {path}")
            except Exception as e:
                print(f"Error loading pretrained weights from {path}: {e}")
        ch = [512, 320, 128, 64]
        d = _Holder()
        d.dec4 = _cfam_block(ch[0])
        d.eucb3 = _eucb(ch[0], ch[1])
        d.skip_enhancer3 = _skip_enhancer(ch[1], num_heads[0])
        d.dec3 = _cfam_block(ch[1])
        d.eucb2 = _eucb(ch[1], ch[2])
        d.skip_enhancer2 = _skip_enhancer(ch[2], num_heads[1])
        d.dec2 = _cfam_block(ch[2])
        d.eucb1 = _eucb(ch[2], ch[3])
        d.skip_enhancer1 = _skip_enhancer(ch[3], num_heads[2])
        d.dec1 = _cfam_block(ch[3])
        self.decoder = d
        fine = [ch[-1] // 2, ch[-1]]
        self.enc = _Seq(_0=_res_block(input_channels, fine[0], 3))
        self.up = _Seq(_1=_res_block(fine[1], fine[0], 3))
        self.rb = _res_block(fine[1], fine[1], 3)
        ob = _Holder()
        ob.conv = _conv_only(fine[1], num_classes, 1, bias=True)
        ob.apply(_init_unet)
        self.out = ob
        self.cfg = dict(input_channels=input_channels, num_classes=num_classes, scale_factors=list(scale_factors),
                        diffatt_num_heads=list(num_heads), dec_up_block="eucb", out_up_block="org")
        self._engines = {}

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        return d

    def _replicate_for_data_parallel(self):
        raise RuntimeError("cenet_b200.CENetOrg does not support nn.DataParallel replicas: run one process per GPU")

    def _engine(self, x, precision=None):
        from ..engine import Engine
        from ..engine_org import EngineOrg
        key = (x.device, precision or Engine.default_precision())
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = EngineOrg(self, x.device, key[1])
        return eng

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("cenet_b200.CENetOrg has no CPU path: move the module and the input to a B200 (`.cuda()`)")
        if self.training:
            raise NotImplementedError("cenet_b200.CENetOrg is an inference path (the reference only evaluates this variant: "
                                      "scripts/synapse.sh TEST_ORG); call .eval() -- training is built for networks.CENet")
        return self._engine(x).forward(x)

    @torch.no_grad()
    def predict(self, x, out=None):
        """fused `argmax(softmax(logits, 1), 1)` -> int64 [B,H,W]"""
        return self._engine(x).forward(x, labels=True, out=out)
