"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch fp32) of the reference's CENet hot path.

This file is the parity ORACLE for the CUDA kernels in cenet_b200/csrc.  It is imported only by tests/,
`__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs.  The product package never
imports it and has no CPU path.

Pinning: tests/golden/make_golden.py imports the real reference (`/root/reference/src/networks`, through
oracle/ref_shim.py) in the build container, loads identical weights and stores its outputs under
tests/golden/*.pt; tests/test_oracle_golden.py checks every function below against those vectors.  The
reference itself ships no tests or golden vectors.

Style: purely functional.  `sd` is a flat state_dict with the reference's key names; `p` is a key prefix.
All citations are relative to /root/reference/src/networks/cenet/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class Cfg:
    """Constructor arguments of the reference CENet (net.py:9-22) that change arithmetic."""
    input_channels: int = 1
    num_classes: int = 1
    scale_factors: List[float] = field(default_factory=lambda: [0.8, 0.4])
    diffatt_num_heads: List[int] = field(default_factory=lambda: [2, 2, 2])
    dec_up_block: str = "eucb"
    out_up_block: str = "eucb"
    skip_mode: str = "cat"                 # DSEBlock: cat([dec, skip]) or dec + skip (dseb.py:155)
    out_merge_mode: str = "cat"            # OutHead.merge (out.py:58-64)
    encoder: str = "pvt_v2_b2"
    # pvt_v2_b2 (pvtv2.py:400-406); b1 / b3 / b4 / b5 differ in depths and MLP ratios only (pvtv2.py:392-431)
    embed_dims: tuple = (64, 128, 320, 512)
    enc_heads: tuple = (1, 2, 5, 8)
    mlp_ratios: tuple = (8, 8, 4, 4)
    depths: tuple = (3, 4, 6, 3)
    sr_ratios: tuple = (8, 4, 2, 1)

    def __post_init__(self):
        table = {"pvt_v2_b1": ((2, 2, 2, 2), (8, 8, 4, 4)), "pvt_v2_b2": ((3, 4, 6, 3), (8, 8, 4, 4)),
                 "pvt_v2_b3": ((3, 4, 18, 3), (8, 8, 4, 4)), "pvt_v2_b4": ((3, 8, 27, 3), (8, 8, 4, 4)),
                 "pvt_v2_b5": ((3, 6, 40, 3), (4, 4, 4, 4))}
        self.depths, self.mlp_ratios = table[self.encoder]


MCA_RATES = {64: (2, 3, 5), 128: (1, 2, 4), 320: (1, 2, 3), 512: (1, 2, 2)}  # decoders.py:64


def lambda_init(depth: int) -> float:
    """multihead_diffattn.py:28-29"""
    return 0.8 - 0.6 * math.exp(-0.3 * depth)


# ------------------------------------------------------------------------------------------------ norms
_STATS_SINK = None        # dict while cenet_forward(..., new_stats=dict) runs in train mode


def _bn(sd, p, x, training=False, eps=1e-5, momentum=0.1):
    """nn.BatchNorm{1,2}d forward (A13): eval -> running stats; train -> biased batch stats, and the running statistics the
    module would hold afterwards (momentum 0.1, UNBIASED batch variance, num_batches_tracked + 1) go to the stats sink."""
    if training:
        if _STATS_SINK is not None:
            with torch.no_grad():
                dims = [d for d in range(x.dim()) if d != 1]
                n = x.numel() // x.shape[1]
                mean = x.mean(dims)
                var_u = x.var(dims, unbiased=False) * (n / max(n - 1, 1))
                _STATS_SINK[p + ".running_mean"] = (1 - momentum) * sd[p + ".running_mean"] + momentum * mean
                _STATS_SINK[p + ".running_var"] = (1 - momentum) * sd[p + ".running_var"] + momentum * var_u
                _STATS_SINK[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def _ln(sd, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ------------------------------------------------------------------------------------------------ encoder
def sr_attention(sd, p, x, H, W, heads, sr):
    """pvtv2.py:88-109 (A11).  x: [B,N,C] tokens."""
    B, N, C = x.shape
    hd = C // heads
    q = _lin(sd, p + ".q", x).view(B, N, heads, hd).transpose(1, 2)
    if sr > 1:
        xi = x.transpose(1, 2).reshape(B, C, H, W)
        xi = F.conv2d(xi, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr)
        xi = xi.flatten(2).transpose(1, 2)
        xi = _ln(sd, p + ".norm", xi, 1e-5)
    else:
        xi = x
    kv = _lin(sd, p + ".kv", xi)
    M = kv.shape[1]
    k = kv[..., :C].view(B, M, heads, hd).transpose(1, 2)
    v = kv[..., C:].view(B, M, heads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(sd, p + ".proj", o)


def mix_ffn(sd, p, x, H, W):
    """pvtv2.py:40-47, 364-370 (A12)."""
    B, N, C = x.shape
    h = _lin(sd, p + ".fc1", x)
    hc = h.shape[-1]
    hi = h.transpose(1, 2).reshape(B, hc, H, W)
    hi = F.conv2d(hi, sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"], padding=1, groups=hc)
    h = F.gelu(hi.flatten(2).transpose(1, 2))
    return _lin(sd, p + ".fc2", h)


def pvt_block(sd, p, x, H, W, heads, sr, drop_mask1=None, drop_mask2=None):
    """pvtv2.py:145-149.  drop_mask*: optional per-sample DropPath multipliers [B,1,1] (A14)."""
    a = sr_attention(sd, p + ".attn", _ln(sd, p + ".norm1", x, 1e-6), H, W, heads, sr)
    x = x + (a if drop_mask1 is None else a * drop_mask1)
    m = mix_ffn(sd, p + ".mlp", _ln(sd, p + ".norm2", x, 1e-6), H, W)
    return x + (m if drop_mask2 is None else m * drop_mask2)


def patch_embed(sd, p, x, k, s):
    """pvtv2.py:185-191: conv(k, stride s, pad k//2) -> tokens -> LayerNorm(1e-5)."""
    y = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], stride=s, padding=k // 2)
    H, W = y.shape[2:]
    return _ln(sd, p + ".norm", y.flatten(2).transpose(1, 2), 1e-5), H, W


def encoder(sd, cfg: Cfg, x, taps=None, drop_masks=None):
    """pvtv2.py:312-348.  Returns the four NCHW pyramid maps.
    drop_masks: optional [2*n_blocks, B] DropPath multipliers (0 or 1/keep; timm drop_path, pvtv2.py:146-147): row 2i scales the
    attention branch of block i, row 2i+1 its MLP branch -- the draws a train-mode step used, so that it can be checked exactly."""
    outs = []
    B = x.shape[0]
    bi = 0
    for s in range(4):
        k, st = (7, 4) if s == 0 else (3, 2)
        t, H, W = patch_embed(sd, f"backbone.patch_embed{s+1}", x, k, st)
        for i in range(cfg.depths[s]):
            m1 = m2 = None
            if drop_masks is not None:
                m1, m2 = drop_masks[2 * bi].view(B, 1, 1), drop_masks[2 * bi + 1].view(B, 1, 1)
            bi += 1
            t = pvt_block(sd, f"backbone.block{s+1}.{i}", t, H, W, cfg.enc_heads[s], cfg.sr_ratios[s], m1, m2)
        t = _ln(sd, f"backbone.norm{s+1}", t, 1e-6)
        x = t.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
        if taps is not None:
            taps[f"backbone.stage{s+1}"] = x
    return outs


# ------------------------------------------------------------------------------------------------ DSEB
def fea(sd, p, x, scale_factors):
    """dseb.py:63-76 + 40-50 (A2): x + w * mean_{i<j} | |x-up(down_i x)| - |x-up(down_j x)| |."""
    H, W = x.shape[2:]
    e = []
    for s in scale_factors:
        d = F.interpolate(x, scale_factor=s, mode="bilinear")
        e.append((x - F.interpolate(d, size=(H, W), mode="bilinear")).abs())
    n = len(e)
    m = n * (n - 1) // 2
    edge = 0
    for i in range(n):
        for j in range(i + 1, n):
            edge = edge + (e[i] - e[j]).abs() / m
    return x + sd[p + ".w"] * edge


def rmsnorm(x, eps):
    """rms_norm.py:15-22 (no affine)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)


def diff_attention(sd, p, t, heads, depth):
    """multihead_diffattn.py:70-129 (A3).  t: [B,N,E] tokens."""
    B, N, E = t.shape
    hd = E // heads // 2
    q = F.linear(t, sd[p + ".q_proj.weight"]).view(B, N, 2 * heads, hd).transpose(1, 2) * hd ** -0.5
    k = F.linear(t, sd[p + ".k_proj.weight"]).view(B, N, 2 * heads, hd).transpose(1, 2)
    v = F.linear(t, sd[p + ".v_proj.weight"]).view(B, N, heads, 2 * hd).transpose(1, 2)
    s = torch.softmax(torch.nan_to_num(q @ k.transpose(-1, -2)), dim=-1)
    li = lambda_init(depth)
    lam = (torch.exp((sd[p + ".lambda_q1"] * sd[p + ".lambda_k1"]).sum())
           - torch.exp((sd[p + ".lambda_q2"] * sd[p + ".lambda_k2"]).sum()) + li)
    s = s.view(B, heads, 2, N, N)
    a = s[:, :, 0] - lam * s[:, :, 1]
    o = rmsnorm(a @ v, 1e-5) * (1.0 - li)
    o = o.transpose(1, 2).reshape(B, N, E)
    return F.linear(o, sd[p + ".out_proj.weight"])


def dse_block(sd, p, skip, dec, scale_factors, heads, depth, mode="cat"):
    """dseb.py:153-165 + 114-118 (A4), use_command='dat-fea'; mode 'cat' or 'add' (dseb.py:155)."""
    y = (dec + skip) if mode == "add" else torch.cat([dec, skip], 1).contiguous()
    B, C2, H, W = y.shape
    x_fea = fea(sd, p + ".boundary", y, scale_factors) + y
    tok = y.view(B, H * W, C2)                                  # pure reinterpretation of the CHW buffer
    gate = diff_attention(sd, p + ".diffattn", tok, heads, depth).reshape(B, C2, H, W)
    z = x_fea + gate * y
    return F.conv2d(z, sd[p + ".mixer.weight"]) + skip


# ------------------------------------------------------------------------------------------------ CFAM
def ccu(sd, p, x, training=False):
    """cfam.py:251-264 (A5)."""
    B, C = x.shape[:2]
    f = x.flatten(2)
    u = torch.stack([f.max(2)[0], f.mean(2), f.std(2, unbiased=False)], -1)     # [B,C,3]
    z = F.conv1d(u, sd[p + ".fc1.weight"], groups=C)
    z = F.conv1d(F.relu(z), sd[p + ".fc2.weight"], groups=C).view(B, C)
    if B > 1:
        z = _bn(sd, p + ".bn", z, training)
    return x * torch.sigmoid(z)[:, :, None, None]


def sep_conv_bn(sd, p, x, rate, training=False):
    """blocks.py:169-185 with depth_activation=True, eps 1e-5 (cfam.py:196-206)."""
    C = x.shape[1]
    x = F.conv2d(x, sd[p + ".depthwise.weight"], padding=rate, dilation=rate, groups=C)
    x = F.relu(_bn(sd, p + ".depthwise_bn", x, training))
    x = F.conv2d(x, sd[p + ".pointwise.weight"])
    return F.relu(_bn(sd, p + ".pointwise_bn", x, training))


def channel_slices(C):
    """cfam.py:178-190 with the hard-coded split [5,5,5,1]."""
    a = int(5 / 16 * C)
    r = int(1 / 16 * C)
    return [(0, a), (a, 2 * a), (2 * a, 3 * a), (3 * a, 3 * a + r)]


def multi_order_dwconv(sd, p, x, training=False):
    """cfam.py:227-241 (A8)."""
    C, H, W = x.shape[1:]
    sl = channel_slices(C)
    outs = []
    for i, r in enumerate(MCA_RATES[C]):
        outs.append(sep_conv_bn(sd, f"{p}.dlps.{i}", x[:, sl[i][0]:sl[i][1]], r, training))
    y = F.adaptive_avg_pool2d(x[:, sl[3][0]:sl[3][1]], (7, 7))
    y = F.conv2d(y, sd[p + ".dlps.3.1.weight"])
    y = F.leaky_relu(_bn(sd, p + ".dlps.3.2", y, training), 0.01)
    y = F.interpolate(y, scale_factor=7, mode="bilinear", align_corners=True)
    if y.shape[2] != H or y.shape[3] != W:
        y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False)
    outs.append(y)
    return F.conv2d(torch.cat(outs, 1), sd[p + ".PW_conv.weight"], sd[p + ".PW_conv.bias"])


def nonlocal_block(sd, p, x, training=False):
    """nlb.py:102-148 (A9)."""
    B, C, H, W = x.shape
    th = F.conv2d(x, sd[p + ".conv_theta.weight"], sd[p + ".conv_theta.bias"]).flatten(2)
    ph = F.conv2d(x, sd[p + ".conv_phi.weight"], sd[p + ".conv_phi.bias"]).flatten(2)
    g = F.conv2d(x, sd[p + ".conv_g.weight"], sd[p + ".conv_g.bias"]).flatten(2)
    a = torch.softmax(th.transpose(1, 2) @ ph * C ** -0.5, dim=2)               # [B,HW,HW]
    y = (g @ a.transpose(1, 2)).view(B, C, H, W)
    pz = _bn(sd, p + ".bn", F.conv2d(y, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"]), training)
    w = sd[p + ".w"]
    return (1 - w) * x + w * pz


def mca(sd, p, x, training=False):
    """cfam.py:298-306 (A7)."""
    x1 = ccu(sd, p + ".ccu", x, training)
    g = F.conv2d(x1, sd[p + ".gate.weight"], sd[p + ".gate.bias"])
    v = multi_order_dwconv(sd, p + ".value", x1, training)
    y = F.conv2d(F.silu(g) * F.silu(v), sd[p + ".proj_2.weight"], sd[p + ".proj_2.bias"]) + x
    return nonlocal_block(sd, p + ".denoising_module", y, training)


def srm(sd, p, x, training=False):
    """cfam.py:93-101 (A6)."""
    u = torch.cat([x.max(1, keepdim=True)[0], x.mean(1, keepdim=True), x.std(1, keepdim=True)], 1)
    f = F.gelu(F.conv2d(u, sd[p + ".pwc.weight"]) + F.conv2d(u, sd[p + ".dwc.weight"], padding=1))
    return x * torch.sigmoid(_bn(sd, p + ".bn", f, training))


def cfam_mlp(sd, p, x, training=False):
    """cfam.py:149-159 (A10)."""
    h = F.conv2d(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])
    h = F.gelu(F.conv2d(h, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=1, groups=h.shape[1]))
    h = srm(sd, p + ".srm", h, training)
    return F.conv2d(h, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])


def cfa_module(sd, p, x, training=False):
    """cfam.py:365-374."""
    x = x + sd[p + ".layer_scale_1"] * mca(sd, p + ".mca", _bn(sd, p + ".norm1", x, training), training)
    return x + sd[p + ".layer_scale_2"] * cfam_mlp(sd, p + ".mlp", _bn(sd, p + ".norm2", x, training), training)


def eucb(sd, p, x, training=False):
    """blocks.py:317-321: nearest x2 -> dw3x3 -> BN -> LeakyReLU(0.2) -> (identity shuffle) -> 1x1."""
    C = x.shape[1]
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, sd[p + ".up_dwc.1.weight"], padding=1, groups=C)
    x = F.leaky_relu(_bn(sd, p + ".up_dwc.2", x, training), 0.2)
    return F.conv2d(x, sd[p + ".pwc.0.weight"], sd[p + ".pwc.0.bias"])


def up_conv(sd, p, x, training=False):
    """blocks.py:206-221: bilinear x2 (align_corners=True) -> 3x3 -> BN -> LeakyReLU(0.2)."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.conv2d(x, sd[p + ".up.1.weight"], padding=1)
    return F.leaky_relu(_bn(sd, p + ".up.2", x, training), 0.2)


def up_rb(sd, p, x, training=False):
    """blocks.py:188-204: bilinear x2 (align_corners=True) -> UnetResBlock(k=3)."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return unet_res_block(sd, p + ".up.1", x, 3, training)


def up_tconv(sd, p, x, training=False):
    """blocks.py:223-243: ConvTranspose2d(k, stride 2, padding (k-1)//2, output_padding 2p+2-k), no bias / norm / activation."""
    w = sd[p + ".up.conv.weight"]
    k = w.shape[-1]
    pad = (k - 2 + 1) // 2
    return F.conv_transpose2d(x, w, stride=2, padding=pad, output_padding=2 * pad + 2 - k)


def _up_block(kind):
    if kind == "eucb":
        return eucb
    if kind == "upcn":
        return up_conv
    if kind == "uprb":
        return up_rb
    if kind == "uptc":
        return up_tconv
    raise NotImplementedError(f"up block '{kind}' is outside the BASELINE configs (SURVEY section 2 row 9)")


def decoder(sd, cfg: Cfg, x4, skips, training=False, taps=None):
    """decoders.py:90-105."""
    up = _up_block(cfg.dec_up_block)
    d = cfa_module(sd, "decoder.dec4", x4, training)
    if taps is not None:
        taps["decoder.dec4"] = d
    for lvl, skip, depth, hi in ((3, skips[0], 4, 0), (2, skips[1], 3, 1), (1, skips[2], 2, 2)):
        d = up(sd, f"decoder.up{lvl}", d, training)
        s = dse_block(sd, f"decoder.skip_enhancer{lvl}", skip, d, cfg.scale_factors, cfg.diffatt_num_heads[hi],
                      depth, cfg.skip_mode.lower())
        if taps is not None:
            taps[f"decoder.up{lvl}"] = d
            taps[f"decoder.skip_enhancer{lvl}"] = s
        d = cfa_module(sd, f"decoder.dec{lvl}", d + s, training)
        if taps is not None:
            taps[f"decoder.dec{lvl}"] = d
    return d


# ------------------------------------------------------------------------------------------------ head
def unet_res_block(sd, p, x, k, training=False):
    """modules/unet.py:201-214, LeakyReLU(0.01)."""
    o = F.conv2d(x, sd[p + ".conv1.conv.weight"], padding=k // 2)
    o = F.leaky_relu(_bn(sd, p + ".norm1", o, training), 0.01)
    o = _bn(sd, p + ".norm2", F.conv2d(o, sd[p + ".conv2.conv.weight"], padding=k // 2), training)
    r = x
    if (p + ".conv3.conv.weight") in sd:
        r = _bn(sd, p + ".norm3", F.conv2d(x, sd[p + ".conv3.conv.weight"]), training)
    return F.leaky_relu(o + r, 0.01)


def out_head(sd, cfg: Cfg, dec, x, training=False, taps=None):
    """out.py:69-75."""
    rb = sd["out.w"] * F.max_pool2d(unet_res_block(sd, "out.rb.0", x, 5, training), 2)
    d = _up_block(cfg.out_up_block)(sd, "out.up", dec, training)
    z = (d + rb) if cfg.out_merge_mode == "add" else torch.cat([d, rb], 1)
    y = unet_res_block(sd, "out.out.0", z, 3, training)
    y = F.conv2d(y, sd["out.out.1.conv.conv.weight"], sd["out.out.1.conv.conv.bias"])
    if taps is not None:
        taps["out.rb"] = rb
        taps["out.up"] = d
        taps["out.pre"] = y
    return F.interpolate(y, scale_factor=2, mode="bilinear")


def cenet_forward(sd: Dict[str, Tensor], cfg: Cfg, x: Tensor, training=False, taps: Optional[dict] = None,
                  new_stats: Optional[dict] = None, drop_masks: Optional[Tensor] = None):
    """new_stats: dict that receives the BatchNorm buffers after one train-mode forward (only with training=True)"""
    global _STATS_SINK
    _STATS_SINK = new_stats if training else None
    try:
        return _cenet_forward(sd, cfg, x, training, taps, drop_masks)
    finally:
        _STATS_SINK = None


def _cenet_forward(sd: Dict[str, Tensor], cfg: Cfg, x: Tensor, training=False, taps: Optional[dict] = None, drop_masks=None):
    """net.py:53-64.  x: [B,Cin,H,W] fp32 -> logits [B,ncls,H,W]."""
    y = torch.cat([x, x, x], 1) if x.shape[1] == 1 else x
    x1, x2, x3, x4 = encoder(sd, cfg, y, taps, drop_masks)
    d = decoder(sd, cfg, x4, [x3, x2, x1], training, taps)
    return out_head(sd, cfg, d, x, training, taps)


def predict_labels(logits: Tensor) -> Tensor:
    """metrics_eval.py:52 / utils_synapse.py:68: argmax(softmax(logits,1),1) -> int64 labels."""
    return torch.argmax(torch.softmax(logits, 1), 1)


# ------------------------------------------------------------------------------------------------ losses
def dice_loss(logits, target, n_classes):
    """utils/core.py:57-80 with softmax=True (A15)."""
    p = torch.softmax(logits, 1)
    loss = 0.0
    for i in range(n_classes):
        t = (target == i).float()
        inter = (p[:, i] * t).sum()
        loss = loss + (1 - (2 * inter + 1e-5) / ((p[:, i] * p[:, i]).sum() + (t * t).sum() + 1e-5))
    return loss / n_classes


def criterion_dice_ce(logits, target, n_classes, w_dice=0.5, w_ce=0.5):
    """utils/core.py:179-188 with --loss_type dice,ce."""
    return w_dice * dice_loss(logits, target, n_classes) + w_ce * F.cross_entropy(logits, target.long())


def boundary_counts(target, n_classes):
    """Integer statistics of BoundaryDoULoss._adaptive_size (utils/core.py:96-107): per class i, S_i = #pixels of class i and
    C_i = #pixels of class i whose 3x3-cross sum over the zero-padded one-hot map is not 5 (a 4-neighbour is another class
    or outside the image).  Returns (C [n_classes], S [n_classes]) as int64."""
    kernel = torch.tensor([[0., 1., 0.], [1., 1., 1.], [0., 1., 0.]]).view(1, 1, 3, 3)
    C, S = [], []
    for i in range(n_classes):
        t = (target == i).float()                                   # [B,H,W]
        y = F.conv2d(t.unsqueeze(1), kernel, padding=1).squeeze(1) * t
        y = torch.where(y == 5, torch.zeros_like(y), y)
        C.append(torch.count_nonzero(y))
        S.append(torch.count_nonzero(t))
    return torch.stack(C), torch.stack(S)


def boundary_dou_loss(logits, target, n_classes):
    """utils/core.py:83-131 (BoundaryDoULoss.forward + _adaptive_size), batch-vectorised; alpha depends on the labels only."""
    p = torch.softmax(logits, 1)
    C, S = boundary_counts(target, n_classes)
    smooth = 1e-5
    loss = 0.0
    for i in range(n_classes):
        t = (target == i).float()
        alpha = 1 - (C[i] + smooth) / (S[i] + smooth)
        alpha = min(float(2 * alpha - 1), 0.8)
        inter = (p[:, i] * t).sum()
        y_sum = (t * t).sum()
        z_sum = (p[:, i] * p[:, i]).sum()
        loss = loss + (z_sum + y_sum - 2 * inter + smooth) / (z_sum + y_sum - (1 + alpha) * inter + smooth)
    return loss / n_classes


def criterion(logits, target, n_classes, w_dice=0.0, w_ce=0.0, w_boundary=0.0):
    """utils/core.py:161-188: weighted sum over --loss_type of dice / ce / boundary."""
    loss = 0.0
    if w_dice:
        loss = loss + w_dice * dice_loss(logits, target, n_classes)
    if w_ce:
        loss = loss + w_ce * F.cross_entropy(logits, target.long())
    if w_boundary:
        loss = loss + w_boundary * boundary_dou_loss(logits, target, n_classes)
    return loss
