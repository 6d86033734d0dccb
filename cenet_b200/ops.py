"""Tensor-level wrappers over the C ABI (one function per kernel family).

Every function takes CUDA tensors, passes raw device pointers + sizes + the CURRENT torch stream to
libcenet_b200.so and returns the output tensor.  Nothing here computes on the host or falls back to torch ops.
Channels-last convention: an activation [B,H,W,C] is also the row-major matrix [B*H*W, C].
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, BF16, F32, GEMM_AUTO, GEMM_MMA, GEMM_SIMT,
                   GEMM_TCGEN05, GemmArgs)

_DT = {torch.float32: F32, torch.bfloat16: BF16}
tag = ""      # free-form region label set by the engine; read only by the profiling proxy (engine._TimedOps)


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"cenet_b200 kernels take float32 / bfloat16 tensors, got {t.dtype}") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cenet_b200 ops need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def _f32(t: Optional[torch.Tensor], what: str):
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
        raise TypeError(f"{what} must be a contiguous float32 tensor")
    return _p(t)


def launch_count() -> int:
    return int(L.load().cenet_launch_count())


# ------------------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldw: int, ldc: int,
         bias=None, bias_per_row=False, row_scale=None, alpha=1.0, act=ACT_NONE, slope=0.0, act_after_res=False,
         res1=None, ldr1=0, res1_cscale=None, res1_scale=1.0, res2=None, ldr2=0, mul=None, ldmul=0, mul_act=ACT_NONE,
         conv=None, batch=1, batch_inner=1, a_bs=(0, 0), w_bs=(0, 0), c_bs=(0, 0), w_nmajor=False, impl=GEMM_AUTO,
         a_off=0, w_off=0, c_off=0, rs_div=1, post_rs=None, post_rs_div=1, a_mmajor=False, r1_off=0, split_ws=None):
    """C[M,N] = epilogue(A[M,K] W[N,K]^T); see cenet_gemm in include/cenet_b200.h.  *_off are element offsets."""
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.batch, g.batch_inner = batch, batch_inner
    g.A, g.a_dtype, g.lda = _p(a) + a_off * a.element_size(), dt(a), lda
    g.a_bs_outer, g.a_bs_inner = a_bs
    if conv is not None:
        g.conv = 1
        (g.Bimg, g.H, g.W, g.Cin, g.KH, g.KW, g.stride, g.pad, g.Ho, g.Wo) = conv
    g.Wt, g.w_dtype, g.ldw = _p(w) + w_off * w.element_size(), dt(w), ldw
    g.w_bs_outer, g.w_bs_inner = w_bs
    g.w_nmajor = int(w_nmajor)
    g.C, g.c_dtype, g.ldc = _p(out) + c_off * out.element_size(), dt(out), ldc
    g.c_bs_outer, g.c_bs_inner = c_bs
    g.alpha = alpha
    g.bias, g.bias_per_row, g.row_scale = _f32(bias, "bias"), int(bias_per_row), _f32(row_scale, "row_scale")
    g.act, g.slope, g.act_after_res = act, slope, int(act_after_res)
    if res1 is not None:
        g.res1, g.res1_dtype, g.ldr1 = _p(res1) + r1_off * res1.element_size(), dt(res1), ldr1
        g.res1_cscale, g.res1_scale = _f32(res1_cscale, "res1_cscale"), res1_scale
    if res2 is not None:
        g.res2, g.res2_dtype, g.ldr2 = _p(res2), dt(res2), ldr2
    if mul is not None:
        g.mul, g.mul_dtype, g.ldmul, g.mul_act = _p(mul), dt(mul), ldmul, mul_act
    g.impl = impl
    if split_ws is not None:                                # fp32 scratch that lets the library split long contractions (split-K)
        g.split_ws, g.split_ws_elems = _f32(split_ws, "split_ws"), split_ws.numel()
    g.a_mmajor, g.rs_div, g.post_row_scale, g.post_rs_div = int(a_mmajor), rs_div, _f32(post_rs, "post_rs"), post_rs_div
    L.call("cenet_gemm", C.byref(g), _stream())
    return out


def linear(x2d: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias=None, **kw):
    """Row-major convenience: x2d [M,K] (contiguous rows), w [N,Kp] (K-major, row pitch Kp >= K), out [M,N]."""
    M, K = x2d.shape
    N = out.shape[1]
    return gemm(x2d, w, out, M=M, N=N, K=K, lda=x2d.stride(0), ldw=w.stride(0), ldc=out.stride(0), bias=bias, **kw)


def conv_nhwc(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, ksize: int, stride: int, pad: int, bias=None, **kw):
    """x [B,H,W,Cin], w [Cout, Kp] with K ordered (kh,kw,cin), out [B,Ho,Wo,(ldc)]."""
    B, H, W, Cin = x.shape
    Ho = (H + 2 * pad - ksize) // stride + 1
    Wo = (W + 2 * pad - ksize) // stride + 1
    N = kw.pop("N", w.shape[0])
    ldc = kw.pop("ldc", out.shape[-1])
    return gemm(x, w, out, M=B * Ho * Wo, N=N, K=ksize * ksize * Cin, lda=Cin, ldw=w.stride(0), ldc=ldc, bias=bias,
                conv=(B, H, W, Cin, ksize, ksize, stride, pad, Ho, Wo), **kw)


def zero_insert_tables(H, W, dev):
    """CSR tap tables of the 2x zero-insertion [H,W] -> [2H,2W] (y[2i,2j] = x[i,j], zeros elsewhere) for `resample`: a transposed
    conv with stride 2 is that followed by a stride-1 conv with flipped taps (uptc up block, blocks.py:223-243)"""
    def csr(n):
        start = torch.zeros(2 * n + 1, dtype=torch.int32)
        start[1::2] = torch.arange(1, n + 1, dtype=torch.int32)            # even output rows own one tap, odd rows none
        start[2::2] = torch.arange(1, n + 1, dtype=torch.int32)
        return start.to(dev), torch.arange(n, dtype=torch.int32).to(dev), torch.ones(n, dtype=torch.float32).to(dev)
    hs, hi, hw = csr(H)
    ws, wi, ww = csr(W)
    return dict(hs=hs, hi=hi, hw=hw, ws=ws, wi=wi, ww=ww)


def resample(x, y, B, Hi, Wi, Ho, Wo, C_, tables, ldx=None, x_off=0, ldy=None, y_off=0, acc=False):
    """sparse separable resampling (cenet_resample): out[i,j] = sum over the CSR taps of row i / column j"""
    t = tables
    L.call("cenet_resample", _p(x) + x_off * x.element_size(), dt(x), C_ if ldx is None else ldx, _p(y) + y_off * y.element_size(), dt(y),
           C_ if ldy is None else ldy, B, Hi, Wi, Ho, Wo, C_, _p(t["hs"]), _p(t["hi"]), _f32(t["hw"], "hw"), _p(t["ws"]), _p(t["wi"]),
           _f32(t["ww"], "ww"), int(acc), _stream())


def add_(dst, src, n, acc):
    """dst[:n] (+)= src[:n]  (DSEBlock / OutHead 'add' merge modes, dseb.py:155, out.py:61)"""
    L.call("cenet_add", _p(dst), _p(src), dt(dst), n, int(acc), _stream())


# ------------------------------------------------------------------------------------------------------ norms
def layernorm(x2d, out, gamma, beta, eps):
    rows, Cc = x2d.shape
    L.call("cenet_layernorm", _p(x2d), dt(x2d), _p(out), dt(out), _f32(gamma, "gamma"), _f32(beta, "beta"), rows, Cc,
           eps, _stream())
    return out


def softmax_rows_(x, rows, n, ld):
    L.call("cenet_softmax_rows", _p(x), dt(x), rows, n, ld, _stream())
    return x


def row_stats(x2d, stats, unbiased=True):
    rows, Cc = x2d.shape
    L.call("cenet_row_stats", _p(x2d), dt(x2d), rows, Cc, x2d.stride(0), int(unbiased), _f32(stats, "stats"), _stream())
    return stats


def rmsnorm_seg(x2d, out, seg, eps, mult):
    rows, Cc = x2d.shape
    L.call("cenet_rmsnorm_seg", _p(x2d), dt(x2d), _p(out), dt(out), rows, Cc, seg, eps, mult, _stream())
    return out


# ------------------------------------------------------------------------------------------------------ dwconv
def mixffn_tail_supported(H, W, Ch, Cc) -> bool:
    return bool(L.load().cenet_mixffn_tail_supported(H, W, Ch, Cc))


def mixffn_tail(h, t, w9c, dw_bias, w2, b2, B, H, W, Ch, Cc):
    """t (fp32 residual stream, in place) += fc2(GELU(dwconv3x3(h) + dw_bias)) + b2 -- one tcgen05 kernel (mixffn_tc.cu)"""
    if h.dtype != torch.bfloat16 or w2.dtype != torch.bfloat16 or t.dtype != torch.float32:
        raise TypeError("mixffn_tail: bf16 hidden / weight and an fp32 residual stream")
    if w2.stride(0) != Ch or w2.shape[0] < Cc:
        raise ValueError("mixffn_tail: w2 must be [C, Ch] with contiguous rows")
    L.call("cenet_mixffn_tail", _p(h), _p(t), _f32(w9c, "w9c"), _f32(dw_bias, "dw_bias"), _p(w2), _f32(b2, "b2"), B, H, W, Ch, Cc,
           _stream())
    return t


def dwconv3x3(x, out, w9c, B, H, W, Cc, *, ldx=None, ldy=None, x_off=0, y_off=0, bias=None, scale=None, shift=None,
              dil=1, up2=False, act=ACT_NONE, slope=0.0, zout=None):
    """H, W are OUTPUT sizes; with up2 the input is [B,H/2,W/2,C].  zout (training): contiguous [B,H,W,C] buffer that
    receives the pre-activation value."""
    ldx = Cc if ldx is None else ldx
    ldy = Cc if ldy is None else ldy
    if zout is not None:
        if scale is not None or up2:
            raise ValueError("dwconv3x3(zout=...) does not take scale/shift/up2")
        L.call("cenet_dwconv3x3_train", _p(x) + x_off * x.element_size(), dt(x), ldx, _p(out) + y_off * out.element_size(),
               dt(out), ldy, _p(zout), _f32(w9c, "w9c"), _f32(bias, "bias"), B, H, W, Cc, dil, act, dt(zout), slope, _stream())
        return out
    L.call("cenet_dwconv3x3", _p(x) + x_off * x.element_size(), dt(x), ldx, _p(out) + y_off * out.element_size(),
           dt(out), ldy, _f32(w9c, "w9c"), _f32(bias, "bias"), _f32(scale, "scale"), _f32(shift, "shift"), B, H, W, Cc,
           dil, int(up2), act, slope, _stream())
    return out


# ------------------------------------------------------------------------------------------------------ layout
def nhwc_to_nchw(x, out, B, HW, Cc, Ctot, coff, ldx=None):
    L.call("cenet_nhwc_to_nchw", _p(x), dt(x), Cc if ldx is None else ldx, _p(out), dt(out), B, HW, Cc, Ctot, coff,
           _stream())
    return out


def nchw_to_nhwc(x, out, B, HW, Cc, ldy=None):
    L.call("cenet_nchw_to_nhwc", _p(x), dt(x), _p(out), dt(out), Cc if ldy is None else ldy, B, HW, Cc, _stream())
    return out


def im2col(x, out, B, H, W, Cin, k, stride, pad, Ho, Wo, Kpad):
    L.call("cenet_im2col", _p(x), dt(x), _p(out), dt(out), B, H, W, Cin, k, k, stride, pad, Ho, Wo, Kpad, _stream())
    return out


def upsample2x_ac(x, out, B, H, W, Cc):
    L.call("cenet_upsample2x_ac", _p(x), dt(x), _p(out), dt(out), B, H, W, Cc, _stream())
    return out


def maxpool2_scale(x, out, ldy, coff, wch, B, H, W, Cc):
    L.call("cenet_maxpool2_scale", _p(x), dt(x), _p(out), dt(out), ldy, coff, _f32(wch, "wch"), B, H, W, Cc, _stream())
    return out


def affine_gate(x, out, scale, shift, gate_bc, B, HW, Cc):
    L.call("cenet_affine_gate", _p(x), dt(x), _p(out), dt(out), _f32(scale, "scale"), _f32(shift, "shift"),
           _f32(gate_bc, "gate"), B, HW, Cc, _stream())
    return out


# ------------------------------------------------------------------------------------------------------ DSEB
def fea_combine(y, gate, z, w_c, B, C2, H, W, scales: Sequence[float]):
    arr = (C.c_float * len(scales))(*[float(s) for s in scales])
    L.call("cenet_fea_combine", _p(y), _p(gate), _p(z), dt(y), _f32(w_c, "fea.w"), B, C2, H, W, arr, len(scales),
           _stream())
    return z


def dog_combine(y, gate, z, w_c, B, C2, H, W, scales: Sequence[float], mode: int):
    """CENetOrg skip enhancer pieces on NCHW planes: mode 1 z = y + w*|y_s0 - y_s1| (DoGEdge), mode 2 z = y + gate*y"""
    sc = list(scales) if scales else [1.0]
    arr = (C.c_float * len(sc))(*[float(v) for v in sc])
    L.call("cenet_dog_combine", _p(y), _p(gate), _p(z), dt(y), _f32(w_c, "dog.w"), B, C2, H, W, arr, len(sc), mode, _stream())
    return z


def diff_combine_(P, npairs, map_elems, lam):
    L.call("cenet_diff_combine", _p(P), dt(P), npairs, map_elems, lam, _stream())
    return P


def diffattn_flash(qkv, out, B, N, E, heads, lam, eps, mult, kmax_ws=None):
    if qkv.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("diffattn_flash is a bf16 kernel")
    L.call("cenet_diffattn_flash", _p(qkv), _p(out), B, N, E, heads, lam, eps, mult, _f32(kmax_ws, "kmax_ws"), _stream())
    return out


def diffattn_flash_padded(qkv, out, B, N, heads, hd_pad, dv_pad, hd_real, lam, eps, mult, kmax_ws=None):
    if qkv.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("diffattn_flash is a bf16 kernel")
    L.call("cenet_diffattn_flash_padded", _p(qkv), _p(out), B, N, heads, hd_pad, dv_pad, hd_real, lam, eps, mult,
           _f32(kmax_ws, "kmax_ws"), _stream())
    return out


def sr_attention(q, kv, out, B, N, Nk, Cc, heads, scale):
    L.call("cenet_sr_attention", _p(q), dt(q), _p(kv), dt(kv), _p(out), dt(out), B, N, Nk, Cc, heads, scale, _stream())
    return out


def nonlocal_flash(tpg, out, B, N, Cc, scale):
    if tpg.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("nonlocal_flash is a bf16 kernel")
    L.call("cenet_nonlocal_flash", _p(tpg), _p(out), B, N, Cc, scale, _stream())
    return out


def attn_tc(q, k, v, o, *, B, heads, Nq, Nk, D, scale, ldq, ldk, ldv, ldo, bq, bk, bv, bo, q_off=0, k_off=0, v_off=0, o_off=0,
            lse=None):
    """tcgen05 flash attention (attn_tc.cu): o = softmax(q k^T * scale) v per (image, head of width D in {64,128}); operands are
    bf16 buffers addressed by element offset / row pitch / image stride, so q, k, v may be column blocks of one tensor."""
    for t in (q, k, v, o):
        if t.dtype != torch.bfloat16:
            raise TypeError("attn_tc is a bf16 kernel")
    a = L.AttnTcArgs()
    a.q, a.k, a.v, a.o = _p(q) + 2 * q_off, _p(k) + 2 * k_off, _p(v) + 2 * v_off, _p(o) + 2 * o_off
    a.lse = _f32(lse, "lse")
    a.ldq, a.ldk, a.ldv, a.ldo, a.bq, a.bk, a.bv, a.bo = ldq, ldk, ldv, ldo, bq, bk, bv, bo
    a.B, a.heads, a.Nq, a.Nk, a.D, a.scale, a.lse_base2 = B, heads, Nq, Nk, D, scale, 0
    import ctypes
    L.call("cenet_attn_tc", ctypes.byref(a), _stream())
    return o


def seg_loss_ws(B, ncls, H, W, device):
    """workspace of `seg_loss` (also large enough for `dice_ce`)"""
    return torch.empty((4 * ncls + 1) * loss_nblocks(B * H * W) + 5 * ncls + 4, device=device, dtype=torch.float32)


def seg_loss(logits, labels, loss_out, dlogits, ws, B, ncls, H, W, w_dice, w_ce, w_boundary, grad_scale=1.0):
    """Criterion (utils/core.py:161-188): w_dice*DiceLoss + w_ce*CrossEntropy + w_boundary*BoundaryDoULoss, fused with its
    gradient.  loss_out [1+ncls] = loss, per-class dice scores."""
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    L.call("cenet_seg_loss", _f32(logits, "logits"), _p(labels), _f32(loss_out, "loss"), _f32(dlogits, "dlogits"),
           _f32(ws, "ws"), B, ncls, H, W, w_dice, w_ce, w_boundary, grad_scale, _stream())
    return loss_out


# ------------------------------------------------------------------------------------------------------ CFAM
def ccu_nchunk(HW: int) -> int:
    return int(L.load().cenet_ccu_nchunk(HW))


def ccu_gate(x, scale, shift, fc1, fc2, bn_scale, bn_shift, gate, ws, B, HW, Cc):
    L.call("cenet_ccu_gate", _p(x), dt(x), _f32(scale, "scale"), _f32(shift, "shift"), _f32(fc1, "fc1"),
           _f32(fc2, "fc2"), _f32(bn_scale, "bn_scale"), _f32(bn_shift, "bn_shift"), _f32(gate, "gate"), _f32(ws, "ws"),
           B, HW, Cc, _stream())
    return gate


def srm_gate(u, gate, pw3, dw27, bn_scale: float, bn_shift: float, B, H, W):
    L.call("cenet_srm_gate", _f32(u, "u"), _f32(gate, "gate"), _f32(pw3, "pw"), _f32(dw27, "dw"), bn_scale, bn_shift, B,
           H, W, _stream())
    return gate


def pool_branch(x, ldx, coff, y, ldy, coff_y, w_rr, bn_scale, bn_shift, slope, pooled_ws, B, H, W, r):
    L.call("cenet_pool_branch", _p(x), dt(x), ldx, coff, _p(y), dt(y), ldy, coff_y, _f32(w_rr, "w"),
           _f32(bn_scale, "bn_scale"), _f32(bn_shift, "bn_shift"), slope, _f32(pooled_ws, "ws"), B, H, W, r, _stream())
    return y


def stem5x5(x, w1, b1, w3, b3, o1, r, B, H, W, Cin, slope):
    L.call("cenet_stem5x5", _p(x), dt(x), _f32(w1, "w1"), _f32(b1, "b1"), _f32(w3, "w3"), _f32(b3, "b3"), _p(o1), _p(r),
           dt(o1), B, H, W, Cin, slope, _stream())
    return o1, r


# ------------------------------------------------------------------------------------------------------ head / loss
def head_upsample_argmax(y, logits, labels, B, h, w, ncls, ldy=0):
    """ldy: pixel pitch of y in elements (0 = ncls)"""
    if labels is not None and labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    L.call("cenet_head_upsample_argmax", _f32(y, "y"), _f32(logits, "logits"), _p(labels), B, h, w, ncls, ldy, _stream())


def volume_labels_counts(pred_patch, iy, ix, label, pred_out, counts, ncls):
    """pred_patch [D,ph,pw] int64 -> pred_out [D,H,W] int64 (nearest gather through iy / ix) and the per-class integer
    counts medpy's `dc` is made of (counts int64 [3*ncls]: intersection | pred size | label size)."""
    D, ph, pw = pred_patch.shape
    H, W = iy.numel(), ix.numel()
    if pred_patch.dtype != torch.int64 or not pred_patch.is_contiguous():
        raise TypeError("pred_patch must be contiguous int64")
    if iy.dtype != torch.int32 or ix.dtype != torch.int32:
        raise TypeError("index tables must be int32")
    kind = None
    if label is not None:
        kind = {torch.float32: 0, torch.int64: 1, torch.uint8: 2}.get(label.dtype)
        if kind is None or not label.is_contiguous() or tuple(label.shape) != (D, H, W):
            raise TypeError("label must be a contiguous float32 / int64 / uint8 tensor of shape [D,H,W]")
    L.call("cenet_volume_labels_counts", _p(pred_patch), ph, pw, _p(iy), _p(ix), _p(label), kind or 0, _p(pred_out), _p(counts),
           D, H, W, ncls, _stream())


def loss_nblocks(npix: int) -> int:
    return int(L.load().cenet_loss_nblocks(npix))


def dice_ce(logits, labels, loss_out, dlogits, ws, B, ncls, HW, w_dice, w_ce, grad_scale=1.0):
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    L.call("cenet_dice_ce", _f32(logits, "logits"), _p(labels), _f32(loss_out, "loss"), _f32(dlogits, "dlogits"),
           _f32(ws, "ws"), B, ncls, HW, w_dice, w_ce, grad_scale, _stream())
    return loss_out
