"""Tensor-level wrappers over the TRAINING entry points of the C ABI (include/cenet_b200.h, section "training").

Same conventions as cenet_b200.ops: CUDA tensors in, raw device pointers + sizes + the current torch stream out; no
host arithmetic, no torch fallback.  `ld*` are row pitches in elements, `*_off` element offsets into the buffer.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_NONE
from .ops import _f32, _p, _stream, dt

ACT_GELU_GRAD = L.ACT_GELU_GRAD


def _po(t, off=0):
    """device pointer of element `off` of tensor t (None stays None)"""
    return None if t is None else _p(t) + off * t.element_size()


def _i32(t, what):
    if t is not None and (t.dtype != torch.int32 or not t.is_contiguous()):
        raise TypeError(f"{what} must be a contiguous int32 tensor")
    return _p(t)


def _ws(ws):
    if ws.dtype != torch.float32:
        raise TypeError("workspace must be float32")
    return _p(ws), ws.numel()


# ------------------------------------------------------------------------------------------------------ GEMM wgrad
def gemm_wgrad(dy, x, dw, *, M, N, K, ldy, y_off, ldx, x_off, T=1, row_scale=None, rs_div=1, dbias=None,
               bias_unscaled=False, ws=None):
    """dw[n, ci, t] = sum_m rs[m] dy[m, n] x[m, t*Cin + ci]   (K = T*Cin);  dbias[n] = sum_m (rs[m]) dy[m, n]"""
    wp, wn = _ws(ws)
    L.call("cenet_gemm_wgrad", _po(dy, y_off), dt(dy), ldy, _po(x, x_off), dt(x), ldx, M, N, K, T, _f32(row_scale, "row_scale"),
           rs_div, _f32(dw, "dw"), _f32(dbias, "dbias"), int(bias_unscaled), wp, wn, _stream())


def gemm_wgrad_partial(dy, x, dw, *, M, N, K, ldy, y_off, ldx, x_off, T=1, row_scale=None, rs_div=1, rs_binary=False,
                       dbias=None, bias_unscaled=False, ws=None):
    """gemm_wgrad with the reduction of its split partials deferred.  Returns the reduction jobs still to run -- a list of
    (src pointer, dst pointer, stride, S, N, K, T, src_ld) for `wgrad_reduce_batch` (empty: dw / dbias are already final) -- and the
    number of workspace floats the partials occupy."""
    wp, wn = _ws(ws)
    S, bp = C.c_int(0), C.c_int(0)
    L.call("cenet_gemm_wgrad_partial", _po(dy, y_off), dt(dy), ldy, _po(x, x_off), dt(x), ldx, M, N, K, T,
           _f32(row_scale, "row_scale"), rs_div, int(rs_binary), _f32(dw, "dw"), _f32(dbias, "dbias"), int(bias_unscaled), wp, wn,
           C.byref(S), C.byref(bp), _stream())
    S, bp = S.value, bp.value
    if S == 0:
        return [], 0
    jobs = [(wp, _p(dw), N * K, S, N, K, T, 0)] if dw is not None else [(wp, None, N * K, S, N, K, T, 0)]
    if bp:
        jobs.append((wp + 4 * S * N * K, _p(dbias), N, S, 1, N, 1, 0))
    return jobs, S * (N * K + (N if bp else 0))


def gemm_wgrad_blocks(dy, x, blocks, *, M, N, K, ldy, y_off, ldx, x_off, ws):
    """weight gradient of a BLOCK-DIAGONAL GEMM: one tcgen05 launch computes all N x K partials, the reduction jobs pick the
    blocks.  blocks: [(dw tensor [rows*cols], row0, rows, col0, cols)].  Returns (jobs, workspace floats used)."""
    jobs, used = gemm_wgrad_partial(dy, x, None, M=M, N=N, K=K, ldy=ldy, y_off=y_off, ldx=ldx, x_off=x_off, ws=ws)
    src, _, stride, S, _, _, _, _ = jobs[0]
    return [(src + 4 * (r0 * K + c0), _p(dw), stride, S, rows, cols, 1, K) for dw, r0, rows, c0, cols in blocks], used


def wgrad_reduce_table(jobs):
    """host side of `wgrad_reduce_batch`: packs the job list (see gemm_wgrad_partial) into the C descriptor array and lays out
    the block ranges; returns (uint8 CPU tensor to be copied to the device, number of jobs, number of blocks)"""
    arr = (L.WgradJob * len(jobs))()
    blk = 0
    lib = L.load()
    for a, (src, dst, stride, S, N, K, T, src_ld) in zip(arr, jobs):
        a.src, a.dst, a.stride, a.S, a.N, a.K, a.T, a.blk0, a.src_ld = src, dst, stride, S, N, K, T, blk, src_ld
        blk += lib.cenet_wgrad_reduce_blocks(C.byref(a))
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone(), len(jobs), blk


def wgrad_reduce_batch(table, njobs, nblocks):
    """table: the device copy of wgrad_reduce_table's tensor"""
    L.call("cenet_wgrad_reduce_batch", _p(table), njobs, nblocks, _stream())


def colsum(x, out, *, rows, C, ld, x_off=0, row_scale=None, rs_div=1, ws=None):
    """out[c] = sum_r row_scale[r // rs_div] * x[r, c]"""
    wp, wn = _ws(ws)
    L.call("cenet_colsum", _po(x, x_off), dt(x), ld, rows, C, _f32(row_scale, "row_scale"), rs_div, _f32(out, "out"), wp, wn, _stream())


def droppath_mask(out, keep, n, B, seed, counter):
    """out [n, B] fp32 = bernoulli(keep[r]) / keep[r]; counter: int64 device tensor [1], advanced by the kernel"""
    L.call("cenet_droppath_mask", _f32(out, "out"), _f32(keep, "keep"), n, B, int(seed) & 0xFFFFFFFFFFFFFFFF, _p(counter), _stream())


def row_scale(x, rs, out, rows, C):
    """out[m, :] = x[m, :] * rs[m]   (contiguous [rows, C])"""
    L.call("cenet_row_scale", _p(x), dt(x), _f32(rs, "rs"), _p(out), rows, C, _stream())


def smallk_dgrad(dy, w, dx, *, rows, K, N, ldw, ldx, acc):
    """dx[m, :N] (+)= dy[m, :K] @ w[:K, :N]   (dy fp32 contiguous, w fp32 with pitch ldw)"""
    L.call("cenet_smallk_dgrad", _f32(dy, "dy"), K, _f32(w, "w"), ldw, _p(dx), dt(dx), ldx, rows, N, int(acc), _stream())


def conv_wgrad(dy, x4, dw, ksize, ws):
    """dw [N,Cin,k,k] of a stride-1 'same' conv: dy [B*H*W, N] contiguous, x4 [B,H,W,Cin] contiguous (no im2col buffer)"""
    B, H, W, Cin = x4.shape
    N = dy.shape[-1]
    wp, wn = _ws(ws)
    L.call("cenet_conv_wgrad", _p(dy), dt(dy), N, _p(x4), dt(x4), Cin, B, H, W, Cin, ksize, N, _f32(dw, "dw"), wp, wn, _stream())


def layernorm_bwd(dy, x, gamma, eps, dx, acc, dgamma, dbeta, ws, defer=False):
    """defer: leave d(gamma) / d(beta) as partial rows in ws and return (reduction jobs for wgrad_reduce_batch, floats used)"""
    rows, Cc = x.shape
    wp, wn = _ws(ws)
    nb = C.c_int(0)
    L.call("cenet_layernorm_bwd", _p(dy), _p(x), dt(x), _f32(gamma, "gamma"), eps, rows, Cc, _p(dx), int(acc),
           _f32(dgamma, "dgamma"), _f32(dbeta, "dbeta"), wp, wn, C.byref(nb) if defer else None, _stream())
    if not defer:
        return [], 0
    S = nb.value
    if S == 0:
        return [], 0
    return [(wp, _p(dgamma), 2 * Cc, S, 1, Cc, 1, 0), (wp + 4 * Cc, _p(dbeta), 2 * Cc, S, 1, Cc, 1, 0)], S * 2 * Cc


# ------------------------------------------------------------------------------------------------------ BatchNorm
def bn_stats(x, rows, C_, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift, mean, rstd, ws, ldx=None, x_off=0,
             frozen=False):
    wp, wn = _ws(ws)
    if nbt.dtype != torch.int64:
        raise TypeError("num_batches_tracked must be int64")
    L.call("cenet_bn_stats", _po(x, x_off), dt(x), C_ if ldx is None else ldx, rows, C_, _f32(gamma, "gamma"),
           _f32(beta, "beta"), _f32(rmean, "running_mean"), _f32(rvar, "running_var"), _p(nbt), momentum, eps,
           int(frozen), _f32(scale, "scale"), _f32(shift, "shift"), _f32(mean, "mean"), _f32(rstd, "rstd"), wp, wn, _stream())


def affine_act(a, out, rows, C_, sa=None, ta=None, b=None, sb=None, tb=None, act=ACT_NONE, slope=0.0, lda=None, a_off=0,
               ldb=None, b_off=0, ldo=None, o_off=0):
    L.call("cenet_affine_act", _po(a, a_off), dt(a), C_ if lda is None else lda, _f32(sa, "sa"), _f32(ta, "ta"),
           _po(b, b_off), dt(b) if b is not None else 0, C_ if ldb is None else ldb, _f32(sb, "sb"), _f32(tb, "tb"),
           _po(out, o_off), dt(out), C_ if ldo is None else ldo, rows, C_, act, slope, _stream())


def bn_bwd(dy, y, a, mean, rstd, gamma, rows, C_, da, dgamma, dbeta, ws, act=ACT_NONE, slope=0.0, acc_da=False, dres=None,
           acc_dres=False, ldy=None, y_off=0, lda=None, a_off=0, lddres=None, dres_off=0, frozen=False):
    wp, wn = _ws(ws)
    ldy = C_ if ldy is None else ldy
    lda = C_ if lda is None else lda
    L.call("cenet_bn_bwd", _po(dy, y_off), dt(dy), _po(y, y_off), dt(y) if y is not None else 0, ldy, _po(a, a_off), dt(a), lda,
           _f32(mean, "mean"), _f32(rstd, "rstd"), _f32(gamma, "gamma"), rows, C_, act, slope, _po(da, a_off), dt(da),
           int(acc_da), _f32(dgamma, "dgamma"), _f32(dbeta, "dbeta"), _po(dres, dres_off),
           dt(dres) if dres is not None else 0, C_ if lddres is None else lddres, int(acc_dres), int(frozen), wp, wn, _stream())


# ------------------------------------------------------------------------------------------------------ depthwise
def dwconv3x3_wgrad(x, dz, dw, dbias, B, H, W, C_, dil, up2, ldx, x_off, ldz, z_off, ws):
    wp, wn = _ws(ws)
    L.call("cenet_dwconv3x3_wgrad", _po(x, x_off), dt(x), ldx, _po(dz, z_off), dt(dz), ldz, B, H, W, C_, dil, int(up2),
           _f32(dw, "dw"), _f32(dbias, "dbias"), wp, wn, _stream())


def sumpool2(full, dx, B, Ho, Wo, C_, acc):
    L.call("cenet_sumpool2", _p(full), dt(full), _p(dx), dt(dx), B, Ho, Wo, C_, int(acc), _stream())


def col2im(dcol, dx, B, H, W, Cin, k, stride, pad, Ho, Wo, Kp, acc):
    L.call("cenet_col2im", _p(dcol), dt(dcol), _p(dx), dt(dx), B, H, W, Cin, k, stride, pad, Ho, Wo, Kp, int(acc), _stream())


# ------------------------------------------------------------------------------------------------------ attention
def _bf16(*ts):
    for t in ts:
        if t.dtype != torch.bfloat16:
            raise TypeError("flash attention kernels are bf16")


def flash_fwd(Q, K, V, O, lse, B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, qo, ldk, ko, ldv, vo, ldo, oo):
    _bf16(Q, K, V, O)
    L.call("cenet_flash_fwd", _po(Q, qo), ldq, _po(K, ko), ldk, _po(V, vo), ldv, _po(O, oo), ldo, _f32(lse, "lse"), B, maps,
           Nq, Nk, dqk, dv, vdiv, scale, _stream())


DIFFATTN_TC_HEAD_DIMS = (8, 16, 32, 64)


def diffattn_fwd_train(qkv, Om, lse, B, N, E, heads, kmax_ws=None):
    """tcgen05 forward of all 2*heads softmax maps of the differential attention (per-map outputs + log2 LSE)"""
    _bf16(qkv, Om)
    L.call("cenet_diffattn_fwd_train", _p(qkv), _p(Om), _f32(lse, "lse"), B, N, E, heads,
           _f32(kmax_ws, "kmax_ws") if kmax_ws is not None else None, _stream())


def flash_bwd(Q, K, V, O, dO, lse, delta, dQ, dK, dV, B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, qo, ldk, ko, ldv, vo, ldo,
              oo, ws=None):
    _bf16(Q, K, V, O, dO, dQ, dK, dV)
    wp, wn = _ws(ws) if ws is not None else (None, 0)
    L.call("cenet_flash_bwd", _po(Q, qo), ldq, _po(K, ko), ldk, _po(V, vo), ldv, _po(O, oo), _po(dO, oo), ldo,
           _f32(lse, "lse"), _f32(delta, "delta"), _po(dQ, qo), _po(dK, ko), _po(dV, vo), B, maps, Nq, Nk, dqk, dv, vdiv,
           scale, wp, wn, _stream())


def softmax_bwd_rows_(P, dP, rows, n):
    L.call("cenet_softmax_bwd_rows", _p(P), _p(dP), dt(P), rows, n, _stream())


def lambda_fwd(lq1, lk1, lq2, lk2, hd, li, lam):
    L.call("cenet_lambda_fwd", _f32(lq1, "lq1"), _f32(lk1, "lk1"), _f32(lq2, "lq2"), _f32(lk2, "lk2"), hd, li,
           _f32(lam, "lam"), _stream())


def lambda_bwd(dlam, lq1, lk1, lq2, lk2, hd, g1, g2, g3, g4):
    L.call("cenet_lambda_bwd", _f32(dlam, "dlam"), _f32(lq1, "lq1"), _f32(lk1, "lk1"), _f32(lq2, "lq2"), _f32(lk2, "lk2"), hd,
           _f32(g1, "g1"), _f32(g2, "g2"), _f32(g3, "g3"), _f32(g4, "g4"), _stream())


def diff_rmsnorm_fwd(Om, lam, o, M, heads, seg, eps, mult):
    L.call("cenet_diff_rmsnorm_fwd", _p(Om), dt(Om), _f32(lam, "lam"), _p(o), M, heads, seg, eps, mult, _stream())


def diff_rmsnorm_bwd(do, Om, lam, dOm, dlam, M, heads, seg, eps, mult, ws):
    wp, wn = _ws(ws)
    L.call("cenet_diff_rmsnorm_bwd", _p(do), _p(Om), dt(Om), _f32(lam, "lam"), _p(dOm), _f32(dlam, "dlam"), M, heads, seg,
           eps, mult, wp, wn, _stream())


# ------------------------------------------------------------------------------------------------------ DSEB
def _fea_bands(mats):
    """[lo, hi) of the non-zero entries of every row and column of the per-axis operators (they are banded).
    Cached ON the operator tensor (an address-keyed cache would hand stale bands to a new tensor that reuses the address)."""
    b = getattr(mats, "_cenet_bands", None)
    if b is None:
        m = mats.detach().float().cpu()
        ns, _, n, _ = m.shape
        out = torch.zeros(ns, 2, 2, n, 2, dtype=torch.int32)
        idx = torch.arange(n)
        for s in range(ns):
            for ax in range(2):
                for t, mat in enumerate((m[s, ax], m[s, ax].t())):
                    nz = mat != 0
                    any_ = nz.any(1)
                    lo = torch.where(nz, idx[None, :], torch.full_like(idx, n)[None, :]).min(1)[0]
                    hi = torch.where(nz, idx[None, :] + 1, torch.zeros_like(idx)[None, :]).max(1)[0]
                    out[s, ax, t, :, 0] = torch.where(any_, lo, torch.zeros_like(lo)).to(torch.int32)
                    out[s, ax, t, :, 1] = torch.where(any_, hi, torch.zeros_like(hi)).to(torch.int32)
        b = out.to(mats.device).contiguous()
        mats._cenet_bands = b
    return b


def fea_bwd(y, gate, dz, w, dy, acc, dgate, dw, B, E, H, W, mats, nscales, ws, ident_mask=0):
    """ident_mask: bit s set when scale factor s is 1.0 (its operator is the identity; the kernel skips its passes)"""
    wp, wn = _ws(ws)
    L.call("cenet_fea_bwd", _p(y), _p(gate), _p(dz), dt(y), _f32(w, "w"), _p(dy), int(acc), _p(dgate), _f32(dw, "dw"), B, E, H,
           W, _f32(mats, "mats"), _i32(_fea_bands(mats), "bands"), mats.shape[-1], nscales, int(ident_mask), wp, wn, _stream())


def nchw_to_nhwc_slice(x, out, B, HW, C_, Ctot, coff, acc):
    L.call("cenet_nchw_to_nhwc_slice", _p(x), dt(x), _p(out), B, HW, C_, Ctot, coff, int(acc), _stream())


def add_(dst, src, n, acc):
    L.call("cenet_add", _p(dst), _p(src), dt(dst), n, int(acc), _stream())


# ------------------------------------------------------------------------------------------------------ CCU
def ccu_stats(xb, u, arg, B, HW, C_, ws):
    wp, wn = _ws(ws)
    L.call("cenet_ccu_stats", _p(xb), dt(xb), _f32(u, "u"), _i32(arg, "arg"), B, HW, C_, wp, wn, _stream())


def ccu_mlp_fwd(u, fc1, fc2, gamma, beta, rmean, rvar, nbt, momentum, eps, gate, save, B, C_, frozen=False):
    L.call("cenet_ccu_mlp_fwd", _f32(u, "u"), _f32(fc1, "fc1"), _f32(fc2, "fc2"), _f32(gamma, "gamma"), _f32(beta, "beta"),
           _f32(rmean, "rmean"), _f32(rvar, "rvar"), _p(nbt), momentum, eps, _f32(gate, "gate"), _f32(save, "save"), B, C_,
           int(frozen), _stream())


def ccu_mlp_bwd(dgate, u, fc1, fc2, gamma, beta, save, du, dfc1, dfc2, dgamma, dbeta, B, C_, frozen=False):
    L.call("cenet_ccu_mlp_bwd", _f32(dgate, "dgate"), _f32(u, "u"), _f32(fc1, "fc1"), _f32(fc2, "fc2"), _f32(gamma, "gamma"),
           _f32(beta, "beta"), _f32(save, "save"), _f32(du, "du"), _f32(dfc1, "dfc1"), _f32(dfc2, "dfc2"), _f32(dgamma, "dgamma"),
           _f32(dbeta, "dbeta"), B, C_, int(frozen), _stream())


def ccu_dgate(dx1, xb, dgate, B, HW, C_, ws):
    wp, wn = _ws(ws)
    L.call("cenet_ccu_dgate", _p(dx1), _p(xb), dt(xb), _f32(dgate, "dgate"), B, HW, C_, wp, wn, _stream())


def ccu_apply_bwd(dx1, xb, gate, u, arg, du, dxb, acc, B, HW, C_):
    L.call("cenet_ccu_apply_bwd", _p(dx1), _p(xb), dt(xb), _f32(gate, "gate"), _f32(u, "u"), _i32(arg, "arg"), _f32(du, "du"),
           _p(dxb), int(acc), B, HW, C_, _stream())


# ------------------------------------------------------------------------------------------------------ SRM
def row_stats_arg(x, u, arg, M, C_):
    L.call("cenet_row_stats_arg", _p(x), dt(x), _f32(u, "u"), _i32(arg, "arg"), M, C_, _stream())


def srm_fwd(u, pw, dw, gamma, beta, rmean, rvar, nbt, momentum, eps, gm, save, st, B, H, W, ws, frozen=False):
    wp, wn = _ws(ws)
    L.call("cenet_srm_fwd", _f32(u, "u"), _f32(pw, "pw"), _f32(dw, "dw"), _f32(gamma, "gamma"), _f32(beta, "beta"),
           _f32(rmean, "rmean"), _f32(rvar, "rvar"), _p(nbt), momentum, eps, _f32(gm, "gm"), _f32(save, "save"), _f32(st, "st"),
           B, H, W, int(frozen), wp, wn, _stream())


def row_dot(a, b, out, M, C_):
    L.call("cenet_row_dot", _p(a), _p(b), dt(a), _f32(out, "out"), M, C_, _stream())


def srm_bwd(dgm, u, gm, save, st, pw, dw, gamma, beta, du, dpw, ddw, dgamma, dbeta, B, H, W, ws, frozen=False):
    wp, wn = _ws(ws)
    L.call("cenet_srm_bwd", _f32(dgm, "dgm"), _f32(u, "u"), _f32(gm, "gm"), _f32(save, "save"), _f32(st, "st"), _f32(pw, "pw"),
           _f32(dw, "dw"), _f32(gamma, "gamma"), _f32(beta, "beta"), _f32(du, "du"), _f32(dpw, "dpw"), _f32(ddw, "ddw"), _f32(dgamma, "dgamma"),
           _f32(dbeta, "dbeta"), B, H, W, int(frozen), wp, wn, _stream())


def srm_apply_bwd(dh3, h2, z, gm, u, arg, du, dz, M, C_):
    L.call("cenet_srm_apply_bwd", _p(dh3), _p(h2), _p(z), dt(h2), _f32(gm, "gm"), _f32(u, "u"), _i32(arg, "arg"),
           _f32(du, "du"), _p(dz), M, C_, _stream())


# ------------------------------------------------------------------------------------------------------ elementwise
def silu_mul_fwd(g, v, out, n):
    L.call("cenet_silu_mul_fwd", _p(g), _p(v), _p(out), dt(g), n, _stream())


def silu_mul_bwd(dout, g, v, dg, dv, n):
    L.call("cenet_silu_mul_bwd", _p(dout), _p(g), _p(v), _p(dg), _p(dv), dt(g), n, _stream())


def ls_combine_fwd(x, y, p, s, t, ls, w, out, M, C_):
    L.call("cenet_ls_combine_fwd", _p(x), _p(y), _p(p), dt(x), _f32(s, "s"), _f32(t, "t"), _f32(ls.reshape(-1), "ls"),
           _f32(w, "w"), _p(out), M, C_, _stream())


def ls_combine_bwd(dout, y, p, s, t, ls, w, dy, acc_dy, dp, dls, dw, M, C_, ws):
    wp, wn = _ws(ws)
    L.call("cenet_ls_combine_bwd", _p(dout), _p(y), _p(p), dt(dout), _f32(s, "s"), _f32(t, "t"), _f32(ls.reshape(-1), "ls"),
           _f32(w, "w"), _p(dy), int(acc_dy), _p(dp), _f32(dls.reshape(-1), "dls"), _f32(dw, "dw"), M, C_, wp, wn, _stream())


# ------------------------------------------------------------------------------------------------------ resampling
def _csr(m):
    """dense [No, Ni] fp32 matrix -> (start int32 [No+1], idx int32 [nnz], w fp32 [nnz])"""
    m = m.float().cpu()
    nz = m != 0
    counts = nz.sum(1)
    start = torch.zeros(m.shape[0] + 1, dtype=torch.int32)
    start[1:] = torch.cumsum(counts, 0).to(torch.int32)
    idx = nz.nonzero()[:, 1].to(torch.int32)
    return start, idx, m[nz]


def gather_cast(src, idx, dst):
    """dst[i] = idx[i] ? src[idx[i]-1] : 0 (cast to dst.dtype); src flat fp32, idx int32, dst flat bf16/fp32"""
    L.call("cenet_gather_cast", _f32(src, "src"), _i32(idx, "map"), _p(dst), dt(dst), dst.numel(), _stream())


def make_tables(Mh, Mw, dev):
    """Sparse per-axis interpolation tables (CSR) for cenet_resample: out[i,j] = sum_h sum_w Mh[i,h] Mw[j,w] in[h,w].
    The matrices come from applying torch's own interpolation / pooling to an identity at plan time, so the tap
    positions and weights are the reference's by construction."""
    hs, hi, hw = _csr(Mh)
    ws_, wi, ww = _csr(Mw)
    return dict(hs=hs.to(dev), hi=hi.to(dev), hw=hw.to(dev), ws=ws_.to(dev), wi=wi.to(dev), ww=ww.to(dev))


def resample(x, y, B, Hi, Wi, Ho, Wo, C_, tables, ldx=None, x_off=0, ldy=None, y_off=0, acc=False):
    t = tables
    L.call("cenet_resample", _po(x, x_off), dt(x), C_ if ldx is None else ldx, _po(y, y_off), dt(y), C_ if ldy is None else ldy,
           B, Hi, Wi, Ho, Wo, C_, _i32(t["hs"], "hs"), _i32(t["hi"], "hi"), _f32(t["hw"], "hw"), _i32(t["ws"], "ws"),
           _i32(t["wi"], "wi"), _f32(t["ww"], "ww"), int(acc), _stream())


def maxpool2_scale_bwd(dz, lddz, coff, rb, w, drb, dw, B, H, W, C_, ws):
    wp, wn = _ws(ws)
    L.call("cenet_maxpool2_scale_bwd", _po(dz, coff), dt(dz), lddz, _p(rb), dt(rb), _f32(w, "w"), _p(drb),
           _f32(dw.reshape(-1), "dw"), B, H, W, C_, wp, wn, _stream())


def head_upsample_bwd(dlogits, dyh, B, h, w, ncls):
    L.call("cenet_head_upsample_bwd", _f32(dlogits, "dlogits"), _f32(dyh, "dyh"), B, h, w, ncls, _stream())


def adamw(p, g, m, v, n, hyper):
    L.call("cenet_adamw", _f32(p, "p"), _f32(g, "g"), _f32(m, "m"), _f32(v, "v"), n, _f32(hyper, "hyper"), _stream())
