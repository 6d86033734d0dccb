"""Launch plan of one CENetOrg forward pass (src/networks/cenet_org/net.py:110-129, decoders.py:112-197) on one B200.

Same kernels, same data layout and same encoder plan as `cenet_b200.engine.Engine` (PVTv2-b2 is byte-identical in the two
variants); this subclass re-wires the decoder and the head:
  gray -> 3 channels  : Conv1x1 + BatchNorm(eval, folded) + ReLU as one K=1 GEMM epilogue, feeding the 7x7 patch embed
  CFAMBlock           : the CFAM plan with dilation rates 6/12/18 and a ReLU image-pooling branch (the state_dict names
                        `attn` / `crm` are mapped onto the `mca` / `ccu` names the shared packing code uses)
  SkipEnhancer        : NCHW island like the DSE block: cat -> DoGEdge (`cenet_dog_combine` mode 1) -> differential attention
                        (depth 1) on the reinterpreted tokens -> `y + gate*y` (mode 2) -> transpose -> 1x1 proj (+bias, +skip, +dec)
  head                : enc = res-block(Cin->32, k3) on the input + max-pool; up = bilinear x2 + res-block(64->32, k3); cat;
                        res-block(64->64, k3); 1x1; fused bilinear x2 + argmax
"""
from __future__ import annotations

import torch

from . import ops
from .engine import Engine
from .ops import ACT_LEAKY, ACT_RELU, GEMM_SIMT


class EngineOrg(Engine):
    def __init__(self, module, device, precision="bf16"):
        super().__init__(module, device, precision)
        self.fixed_rates = (6, 12, 18)          # cenet_org/modules/cfam.py:298
        self.pool_slope = 0.0                   # nn.ReLU in the image-pooling branch (cfam.py:229)
        self.eucb_slope = 0.0                   # EUCB default activation='relu' (decoders.py:84)

    def _sd(self):
        """state dict with the decoder names translated to the ones Engine's packing helpers read"""
        out = {}
        for k, v in self.mod.state_dict().items():
            k2 = k.replace(".attn.crm.", ".attn.ccu.").replace(".attn.", ".mca.") if k.startswith("decoder.dec") else k
            for lvl in (1, 2, 3):
                k2 = k2.replace(f"decoder.eucb{lvl}.", f"decoder.up{lvl}.")
            out[k2] = v.detach().to(self.dev, torch.float32)
        return out

    # ------------------------------------------------------------------------------------------------ packing
    def pack(self):
        sd = self._sd()
        cfg = self.cfg
        P, M = self._put, self._put_mat
        self._pack_encoder(sd, sum_gray=False)                   # the patch embed always sees 3 channels here
        if cfg["input_channels"] == 1:                           # relu(bn(conv1x1(x))) = relu(x * s[c] + t[c])
            s, t = self._bn_fold(sd, "conv.1")
            self._put("gray.w", (sd["conv.0.weight"].reshape(3) * s).reshape(3, 1), self.T)
            self._put("gray.b", sd["conv.0.bias"] * s + t)
        for name, Cc in (("dec4", 512), ("dec3", 320), ("dec2", 128), ("dec1", 64)):
            self._pack_cfam(sd, f"decoder.{name}", Cc)
        for lvl, hi in ((3, 0), (2, 1), (1, 2)):
            self._pack_up(sd, f"decoder.up{lvl}", "eucb")
            p = f"decoder.skip_enhancer{lvl}"
            P(p + ".dog_w", sd[p + ".boundary.w"].reshape(-1))
            self._pack_diffattn(sd, p + ".diffattn", cfg["diffatt_num_heads"][hi], 1)
            M(p + ".proj.w", sd[p + ".proj.weight"].flatten(1))
            P(p + ".proj.b", sd[p + ".proj.bias"])
        # head: three UnetResBlocks (BatchNorm folded into the convs) and the 1x1 classifier
        self._pack_resblock(sd, "enc.0", 3)
        cin = cfg["input_channels"]
        s1, t1 = self._bn_fold(sd, "enc.0.norm1")               # first conv on the raw image: the CUDA-core stem kernel,
        w3x3 = sd["enc.0.conv1.conv.weight"] * s1[:, None, None, None]     # its 3x3 filter embedded in the kernel's 5x5 window
        w5 = torch.zeros(w3x3.shape[0], cin, 5, 5, device=self.dev)
        w5[:, :, 1:4, 1:4] = w3x3
        self._put("enc.0.stem.w1", self._conv_mat(w5)); self._put("enc.0.stem.b1", t1)
        s3, t3 = self._bn_fold(sd, "enc.0.norm3")
        self._put("enc.0.stem.w3", sd["enc.0.conv3.conv.weight"].flatten(1) * s3[:, None]); self._put("enc.0.stem.b3", t3)
        self._pack_resblock(sd, "up.1", 3)
        self._pack_resblock(sd, "rb", 3)
        M("head.w", sd["out.conv.conv.weight"].flatten(1))
        P("head.b", sd["out.conv.conv.bias"])
        self._put("ones32", torch.ones(32))
        self._wver = self._weights_version()
        self._graphs.clear()

    # ------------------------------------------------------------------------------------------------ blocks
    def _skip_enhancer(self, skip, dec, B, H, W, Cc, p, heads, key):
        """decoders.py:138-144 (mode='cat'); returns proj(z) + skip + dec (== dec + SkipEnhancer(skip, dec), decoders.py:186)"""
        w = self.w
        ops.tag = key
        HW, E = H * W, 2 * Cc
        y = self.buf(key + ".y", (B, E, H, W))                         # NCHW cat([dec, skip])
        ops.nhwc_to_nchw(dec, y, B, HW, Cc, E, 0)
        ops.nhwc_to_nchw(skip, y, B, HW, Cc, E, Cc)
        yd = self.buf(key + ".yd", (B, E, H, W))                       # DoGEdge(y)
        ops.dog_combine(y, None, yd, w[p + ".dog_w"], B, E, H, W, self.cfg["scale_factors"], 1)
        tok = yd.view(B * HW, E)                                       # the reference's `.view` reinterpretation
        gate = self._diff_attention(tok, B, HW, E, heads, p + ".diffattn", key + ".da")
        z = self.buf(key + ".z", (B, E, H, W))
        ops.dog_combine(yd, gate, z, w[p + ".dog_w"], B, E, H, W, None, 2)          # z = y' + diffattn(tok) * tok
        zt = self.buf(key + ".zt", (B * HW, E))
        ops.nchw_to_nhwc(z, zt, B, HW, E)
        out = self.buf(key + ".out", (B * HW, Cc))
        ops.linear(zt, w[p + ".proj.w"], out, bias=w[p + ".proj.b"], res1=skip, ldr1=Cc, res2=dec, ldr2=Cc, impl=self.gemm_impl)
        if self.taps is not None:
            self.taps[p] = (out.float() - dec.float()).reshape(B, H, W, Cc).permute(0, 3, 1, 2).clone()
        return out

    def _resblock_to(self, x, B, H, W, Cin, Cout, p, key, out, ldc, c_off):
        """UnetResBlock k=3 (modules/unet.py:201-214, BatchNorm folded) writing its output into columns [c_off, c_off+Cout) of
        `out` (row pitch ldc): conv1 + LeakyReLU(0.01); conv2; residual through conv3 (1x1) when Cin != Cout; LeakyReLU"""
        w = self.w
        ops.tag = key
        Mtok = B * H * W
        x4 = x.view(B, H, W, Cin)
        o1 = self.buf(key + ".o1", (B, H, W, Cout))
        ops.conv_nhwc(x4, w[p + ".c1.w"], o1, 3, 1, 1, bias=w[p + ".c1.b"], act=ACT_LEAKY, slope=0.01, impl=self.gemm_impl)
        if (p + ".c3.w") in w:
            r = self.buf(key + ".r", (Mtok, Cout))
            ops.gemm(x, w[p + ".c3.w"], r, M=Mtok, N=Cout, K=Cin, lda=Cin, ldw=w[p + ".c3.w"].shape[1], ldc=Cout,
                     bias=w[p + ".c3.b"], impl=self.gemm_impl)
        else:
            r = x
        ops.conv_nhwc(o1, w[p + ".c2.w"], out, 3, 1, 1, bias=w[p + ".c2.b"], act=ACT_LEAKY, slope=0.01, act_after_res=True,
                      res1=r, ldr1=Cout, N=Cout, ldc=ldc, c_off=c_off, impl=self.gemm_impl)
        return out

    # ------------------------------------------------------------------------------------------------ forward
    def _run(self, x_in, B, H, W, out_logits, out_labels):
        cfg, w = self.cfg, self.w
        Cin, ncls = cfg["input_channels"], cfg["num_classes"]
        xc = self.buf("x", (B * H * W, Cin))
        ops.tag = "input"
        if Cin == 1:
            ops.affine_gate(x_in, xc, None, None, None, B, H * W, 1)
            y3 = self.buf("x3", (B * H * W, 3))
            ops.gemm(xc, w["gray.w"], y3, M=B * H * W, N=3, K=1, lda=1, ldw=1, ldc=3, bias=w["gray.b"], act=ACT_RELU,
                     impl=GEMM_SIMT)                                   # K = 1: an elementwise affine + ReLU, not a contraction
        else:
            ops.nchw_to_nhwc(x_in, xc, B, H * W, Cin)
            y3 = xc
        feats = self._encoder(y3, B, H, W, 3)
        (x1, H1, W1, C1), (x2, H2, W2, C2), (x3, H3, W3, C3), (x4, H4, W4, C4) = feats
        d = self._cfam(x4, B, H4, W4, C4, "decoder.dec4", "dec4")
        heads = cfg["diffatt_num_heads"]
        for lvl, (sk, Hs, Ws, Cs), hi, Cprev in ((3, feats[2], 0, C4), (2, feats[1], 1, C3), (1, feats[0], 2, C2)):
            up = self._up(d, B, Hs // 2, Ws // 2, Cprev, Cs, f"decoder.up{lvl}", "eucb", f"up{lvl}")
            self._tap(f"decoder.eucb{lvl}", up, B, Hs, Ws, Cs)
            xin = self._skip_enhancer(sk, up, B, Hs, Ws, Cs, f"decoder.skip_enhancer{lvl}", heads[hi], f"se{lvl}")
            d = self._cfam(xin, B, Hs, Ws, Cs, f"decoder.dec{lvl}", f"dec{lvl}")
        # ---- head (net.py:117-127) ----
        om = C1 // 2
        Hh, Wh = H // 2, W // 2
        z = self.buf("head.z", (B * Hh * Wh, 2 * om))                  # cat([dec, enc], 1)
        ops.tag = "head.enc"
        o1 = self.buf("head.enc.o1", (B, H, W, om))
        rres = self.buf("head.enc.r", (B * H * W, om))
        ops.stem5x5(xc, w["enc.0.stem.w1"], w["enc.0.stem.b1"], w["enc.0.stem.w3"], w["enc.0.stem.b3"], o1, rres, B, H, W, Cin, 0.01)
        e2 = self.buf("head.enc.o2", (B * H * W, om))
        ops.conv_nhwc(o1, w["enc.0.c2.w"], e2, 3, 1, 1, bias=w["enc.0.c2.b"], act=ACT_LEAKY, slope=0.01, act_after_res=True,
                      res1=rres, ldr1=om, impl=self.gemm_impl)
        ops.maxpool2_scale(e2, z, 2 * om, om, w["ones32"][:om].contiguous(), B, H, W, om)
        ops.tag = "head.up"
        t = self.buf("head.up.t", (B, Hh, Wh, C1))
        ops.upsample2x_ac(d, t, B, H1, W1, C1)
        self._resblock_to(t.view(B * Hh * Wh, C1), B, Hh, Wh, C1, om, "up.1", "head.up", z, 2 * om, 0)
        o = self._resblock(z, B, Hh, Wh, 2 * om, 2 * om, 3, "rb", "head.rb")
        yh = self.buf("head.y", (B * Hh * Wh, ncls), torch.float32)
        ops.tag = "head.logits"
        ops.linear(o, w["head.w"], yh, bias=w["head.b"], impl=self.gemm_impl)
        ops.head_upsample_argmax(yh, out_logits, out_labels, B, Hh, Wh, ncls)
