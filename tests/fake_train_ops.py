"""TEST-ONLY torch emulation of cenet_b200.train_ops (the C-ABI training kernels): same signatures, same buffer / pitch /
offset semantics, fp32 arithmetic rounded to the buffer dtype on store.

Two uses: (1) `-m "not gpu"`: run `cenet_b200.train.TrainEngine`'s launch plan on CPU and check loss + every parameter
gradient against autograd through the oracle (host logic: tape order, gradient accumulation, slices, packing);
(2) `-m gpu`: per-kernel reference for tests/test_gpu_train_ops.py.  Never imported by the product.
Backward emulations deliberately use torch.autograd on the forward formula -- they are the checker, not the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from fake_ops import ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, _LAUNCHES, _act, _as, _flat

ACT_GELU_GRAD = 6


def _m(t, rows, C, ld=None, off=0):
    """strided [rows, C] view of a contiguous buffer (pitch ld, element offset off)"""
    ld = C if ld is None else ld
    return _as(_flat(t), (rows, C), (ld, 1), off)


def _store(view, val, acc=False):
    if acc:
        view.copy_(view.float() + val)
    else:
        view.copy_(val)


def _act_grad_from_out(y, act, slope):
    if act == ACT_RELU:
        return (y > 0).float()
    if act == ACT_LEAKY:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    assert act == ACT_NONE
    return torch.ones_like(y)


# ------------------------------------------------------------------------------------------------------ GEMM wgrad
def gemm_wgrad(dy, x, dw, *, M, N, K, ldy, y_off, ldx, x_off, T=1, row_scale=None, rs_div=1, dbias=None,
               bias_unscaled=False, ws=None):
    _LAUNCHES[0] += 2
    dyv = _m(dy, M, N, ldy, y_off).float()
    xv = _m(x, M, K, ldx, x_off).float()
    dys = dyv
    if row_scale is not None:
        rs = row_scale.view(-1)[torch.arange(M) // rs_div]
        dys = dyv * rs[:, None]
    g = dys.t() @ xv                                        # [N, K], k = t*Cin + ci
    Cin = K // T
    dw.view(N, Cin, T).copy_(g.view(N, T, Cin).transpose(1, 2))
    if dbias is not None:
        dbias.view(-1)[:N].copy_((dyv if bias_unscaled else dys).sum(0))


def gemm_wgrad_partial(dy, x, dw, *, rs_binary=False, **kw):
    """the emulation has no split partials: the result is final, nothing is left to reduce"""
    gemm_wgrad(dy, x, dw, **kw)
    return [], 0


def gemm_wgrad_blocks(dy, x, blocks, *, M, N, K, ldy, y_off, ldx, x_off, ws):
    full = torch.zeros(N * K)
    gemm_wgrad(dy, x, full, M=M, N=N, K=K, ldy=ldy, y_off=y_off, ldx=ldx, x_off=x_off)
    for dw, r0, rows, c0, cols in blocks:
        dw.view(-1)[:rows * cols].copy_(full.view(N, K)[r0:r0 + rows, c0:c0 + cols].reshape(-1))
    return [], 0


def wgrad_reduce_table(jobs):
    raise AssertionError("the emulation never defers a reduction")


def wgrad_reduce_batch(table, njobs, nblocks):
    raise AssertionError("the emulation never defers a reduction")


def colsum(x, out, *, rows, C, ld, x_off=0, row_scale=None, rs_div=1, ws=None):
    _LAUNCHES[0] += 2
    v = _m(x, rows, C, ld, x_off).float()
    if row_scale is not None:
        v = v * row_scale.view(-1)[torch.arange(rows) // rs_div][:, None]
    out.view(-1)[:C].copy_(v.sum(0))


def droppath_mask(out, keep, n, B, seed, counter):
    _LAUNCHES[0] += 1
    g = torch.Generator().manual_seed(int(seed) % (2 ** 31) + int(counter.item()))
    k = keep.view(n, 1).float().cpu()
    out.copy_((torch.rand(n, B, generator=g) < k).float() / k)
    counter.add_(1)


def row_scale(x, rs, out, rows, C):
    _LAUNCHES[0] += 1
    _flat(out)[:rows * C].view(rows, C).copy_(_flat(x)[:rows * C].view(rows, C).float() * rs.view(-1)[:rows, None])


def smallk_dgrad(dy, w, dx, *, rows, K, N, ldw, ldx, acc):
    _LAUNCHES[0] += 1
    wv = _as(_flat(w), (K, N), (ldw, 1), 0).float()
    v = dy.view(rows, K).float() @ wv
    _store(_m(dx, rows, N, ldx, 0), v, acc)


def conv_wgrad(dy, x4, dw, ksize, ws):
    _LAUNCHES[0] += 2
    B, H, W, Cin = x4.shape
    N = dy.shape[-1]
    w = torch.zeros(N, Cin, ksize, ksize, requires_grad=True)
    out = F.conv2d(x4.float().permute(0, 3, 1, 2), w, padding=ksize // 2)
    dw.copy_(torch.autograd.grad(out, w, _flat(dy).view(B, H, W, N).float().permute(0, 3, 1, 2))[0])


def layernorm_bwd(dy, x, gamma, eps, dx, acc, dgamma, dbeta, ws, defer=False):
    _LAUNCHES[0] += 2
    xf = x.float().detach().requires_grad_(True)
    g = gamma.detach().clone().requires_grad_(True)
    b = torch.zeros_like(g).requires_grad_(True)
    y = F.layer_norm(xf, (xf.shape[-1],), g, b, eps)
    gx, gg, gb = torch.autograd.grad(y, (xf, g, b), dy.float())
    _store(dx, gx, acc)
    dgamma.copy_(gg)
    dbeta.copy_(gb)
    return [], 0


# ------------------------------------------------------------------------------------------------------ BatchNorm
def bn_stats(x, rows, C, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift, mean, rstd, ws, ldx=None, x_off=0,
             frozen=False):
    _LAUNCHES[0] += 2
    if frozen:                                   # eval-mode BatchNorm: running statistics, nothing updated
        r = torch.rsqrt(rvar + eps)
        mean.copy_(rmean); rstd.copy_(r); scale.copy_(gamma * r); shift.copy_(beta - rmean * gamma * r)
        return
    xv = _m(x, rows, C, ldx, x_off).float()
    mu = xv.mean(0)
    var = xv.var(0, unbiased=False)
    r = torch.rsqrt(var + eps)
    mean.copy_(mu)
    rstd.copy_(r)
    scale.copy_(gamma * r)
    shift.copy_(beta - mu * gamma * r)
    rmean.mul_(1 - momentum).add_(momentum * mu)
    rvar.mul_(1 - momentum).add_(momentum * var * (rows / max(rows - 1, 1)))
    nbt.add_(1)


def affine_act(a, out, rows, C, sa=None, ta=None, b=None, sb=None, tb=None, act=ACT_NONE, slope=0.0, lda=None, a_off=0,
               ldb=None, b_off=0, ldo=None, o_off=0):
    _LAUNCHES[0] += 1
    v = _m(a, rows, C, lda, a_off).float()
    if sa is not None:
        v = v * sa + ta
    if b is not None:
        w = _m(b, rows, C, ldb, b_off).float()
        if sb is not None:
            w = w * sb + tb
        v = v + w
    _m(out, rows, C, ldo, o_off).copy_(_act(v, act, slope))


def bn_bwd(dy, y, a, mean, rstd, gamma, rows, C, da, dgamma, dbeta, ws, act=ACT_NONE, slope=0.0, acc_da=False, dres=None,
           acc_dres=False, ldy=None, y_off=0, lda=None, a_off=0, lddres=None, dres_off=0, frozen=False):
    _LAUNCHES[0] += 3
    g = _m(dy, rows, C, ldy, y_off).float()
    if y is not None and act != ACT_NONE:
        g = g * _act_grad_from_out(_m(y, rows, C, ldy, y_off).float(), act, slope)
    xh = (_m(a, rows, C, lda, a_off).float() - mean) * rstd
    db = g.sum(0)
    dg = (g * xh).sum(0)
    dx = gamma * rstd * (g if frozen else (g - db / rows - xh * dg / rows))
    _store(_m(da, rows, C, lda, a_off), dx, acc_da)
    dgamma.copy_(dg)
    dbeta.copy_(db)
    if dres is not None:
        _store(_m(dres, rows, C, lddres, dres_off), g, acc_dres)


# ------------------------------------------------------------------------------------------------------ depthwise
def dwconv3x3_wgrad(x, dz, dw, dbias, B, H, W, C, dil, up2, ldx, x_off, ldz, z_off, ws):
    _LAUNCHES[0] += 2
    Hi, Wi = (H // 2, W // 2) if up2 else (H, W)
    xi = _as(_flat(x), (B, Hi, Wi, C), (Hi * Wi * ldx, Wi * ldx, ldx, 1), x_off).float().permute(0, 3, 1, 2)
    if up2:
        xi = F.interpolate(xi, scale_factor=2, mode="nearest")
    g = _as(_flat(dz), (B, H, W, C), (H * W * ldz, W * ldz, ldz, 1), z_off).float().permute(0, 3, 1, 2)
    w = torch.zeros(C, 1, 3, 3, requires_grad=True)
    out = F.conv2d(xi, w, padding=dil, dilation=dil, groups=C)
    dw.copy_(torch.autograd.grad(out, w, g)[0])
    if dbias is not None:
        dbias.copy_(g.sum((0, 2, 3)))


def sumpool2(full, dx, B, Ho, Wo, C, acc):
    _LAUNCHES[0] += 1
    f = _flat(full).view(B, Ho, 2, Wo, 2, C).float().sum((2, 4))
    _store(_flat(dx).view(B, Ho, Wo, C), f, acc)


def col2im(dcol, dx, B, H, W, Cin, k, stride, pad, Ho, Wo, Kp, acc):
    _LAUNCHES[0] += 1
    c = _flat(dcol).view(B, Ho * Wo, Kp)[:, :, :k * k * Cin].float()
    c = c.view(B, Ho * Wo, k * k, Cin).permute(0, 3, 2, 1).reshape(B, Cin * k * k, Ho * Wo)
    img = F.fold(c, (H, W), k, padding=pad, stride=stride)                    # [B,Cin,H,W]
    _store(_flat(dx).view(B, H, W, Cin), img.permute(0, 2, 3, 1), acc)


# ------------------------------------------------------------------------------------------------------ attention
def _heads(t, B, N, ld, off, maps, d, div=1):
    """[B, maps, N, d] view: head m at columns off + (m // div) * d"""
    v = _as(_flat(t), (B, maps // div, N, d), (N * ld, d, ld, 1), off)
    return v.repeat_interleave(div, 1) if div > 1 else v


def flash_fwd(Q, K, V, O, lse, B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, qo, ldk, ko, ldv, vo, ldo, oo):
    _LAUNCHES[0] += 1
    q = _heads(Q, B, Nq, ldq, qo, maps, dqk).float()
    k = _heads(K, B, Nk, ldk, ko, maps, dqk).float()
    v = _heads(V, B, Nk, ldv, vo, maps, dv, vdiv).float()
    s = q @ k.transpose(-1, -2) * scale
    lse.view(B, maps, Nq).copy_(torch.logsumexp(s, -1))
    _as(_flat(O), (B, maps, Nq, dv), (Nq * ldo, dv, ldo, 1), oo).copy_(torch.softmax(s, -1) @ v)


DIFFATTN_TC_HEAD_DIMS = (8, 16, 32, 64)


def diffattn_fwd_train(qkv, Om, lse, B, N, E, heads, kmax_ws=None):
    hd = E // (2 * heads)
    flash_fwd(qkv, qkv, qkv, Om, lse, B, 2 * heads, N, N, hd, 2 * hd, 2, hd ** -0.5, 3 * E, 0, 3 * E, E, 3 * E, 2 * E, 2 * E, 0)


def flash_bwd(Q, K, V, O, dO, lse, delta, dQ, dK, dV, B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, qo, ldk, ko, ldv, vo, ldo, oo,
              ws=None):
    _LAUNCHES[0] += 3
    q = _heads(Q, B, Nq, ldq, qo, maps, dqk).float().requires_grad_(True)
    k = _heads(K, B, Nk, ldk, ko, maps, dqk).float().requires_grad_(True)
    v0 = _as(_flat(V), (B, maps // vdiv, Nk, dv), (Nk * ldv, dv, ldv, 1), vo).float().requires_grad_(True)
    v = v0.repeat_interleave(vdiv, 1) if vdiv > 1 else v0
    o = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v
    g = _as(_flat(dO), (B, maps, Nq, dv), (Nq * ldo, dv, ldo, 1), oo).float()
    gq, gk, gv = torch.autograd.grad(o, (q, k, v0), g)
    _as(_flat(dQ), (B, maps, Nq, dqk), (Nq * ldq, dqk, ldq, 1), qo).copy_(gq)
    _as(_flat(dK), (B, maps, Nk, dqk), (Nk * ldk, dqk, ldk, 1), ko).copy_(gk)
    _as(_flat(dV), (B, maps // vdiv, Nk, dv), (Nk * ldv, dv, ldv, 1), vo).copy_(gv)


def softmax_bwd_rows_(P, dP, rows, n):
    _LAUNCHES[0] += 1
    p = _flat(P).view(rows, n).float()
    d = _flat(dP).view(rows, n)
    df = d.float()
    d.copy_(p * (df - (p * df).sum(-1, keepdim=True)))


def lambda_fwd(lq1, lk1, lq2, lk2, hd, li, lam):
    _LAUNCHES[0] += 1
    lam[0] = torch.exp((lq1 * lk1).sum()) - torch.exp((lq2 * lk2).sum()) + li


def lambda_bwd(dlam, lq1, lk1, lq2, lk2, hd, g1, g2, g3, g4):
    _LAUNCHES[0] += 1
    e1, e2 = torch.exp((lq1 * lk1).sum()), torch.exp((lq2 * lk2).sum())
    d = dlam[0]
    g1.copy_(d * e1 * lk1)
    g2.copy_(d * e1 * lq1)
    g3.copy_(-d * e2 * lk2)
    g4.copy_(-d * e2 * lq2)


def _diff_rms(Om, lam, M, heads, seg, eps, mult):
    o = Om.view(M, heads, 2, seg)
    a = o[:, :, 0] - lam * o[:, :, 1]
    return (a * torch.rsqrt(a.pow(2).mean(-1, keepdim=True) + eps) * mult).reshape(M, heads * seg)


def diff_rmsnorm_fwd(Om, lam, o, M, heads, seg, eps, mult):
    _LAUNCHES[0] += 1
    o.copy_(_diff_rms(Om.float(), lam[0], M, heads, seg, eps, mult))


def diff_rmsnorm_bwd(do, Om, lam, dOm, dlam, M, heads, seg, eps, mult, ws):
    _LAUNCHES[0] += 2
    om = Om.float().detach().requires_grad_(True)
    l = lam[0].detach().clone().requires_grad_(True)
    y = _diff_rms(om, l, M, heads, seg, eps, mult)
    g1, g2 = torch.autograd.grad(y, (om, l), do.float())
    dOm.copy_(g1)
    dlam[0] = g2


# ------------------------------------------------------------------------------------------------------ DSEB
def fea_bwd(y, gate, dz, w, dy, acc, dgate, dw, B, E, H, W, mats, nscales, ws, ident_mask=0):
    """z = 2y + w*edge(y) + gate*y with the per-axis operators A_s = mats[s, axis, :n, :n] (up(down(.)))"""
    _LAUNCHES[0] += 2
    yf = _flat(y).view(B, E, H, W).float().detach().requires_grad_(True)
    gf = _flat(gate).view(B, E, H, W).float().detach().requires_grad_(True)
    wf = w.detach().clone().view(1, E, 1, 1).requires_grad_(True)
    e = []
    for s in range(nscales):
        Ah, Aw = mats[s, 0, :H, :H].float(), mats[s, 1, :W, :W].float()
        xs = torch.einsum("ih,bchw,jw->bcij", Ah, yf, Aw)
        e.append((yf - xs).abs())
    m = nscales * (nscales - 1) // 2
    edge = 0
    for i in range(nscales):
        for j in range(i + 1, nscales):
            edge = edge + (e[i] - e[j]).abs() / m
    z = 2 * yf + wf * edge + gf * yf
    gy, gg, gw = torch.autograd.grad(z, (yf, gf, wf), _flat(dz).view(B, E, H, W).float())
    _store(_flat(dy).view(B, E, H, W), gy, acc)
    _flat(dgate).view(B, E, H, W).copy_(gg)
    dw.copy_(gw.view(dw.shape))


def nchw_to_nhwc_slice(x, out, B, HW, C, Ctot, coff, acc):
    _LAUNCHES[0] += 1
    xi = _as(_flat(x), (B, C, HW), (Ctot * HW, HW, 1), coff * HW).float()
    _store(_flat(out).view(B, HW, C), xi.transpose(1, 2), acc)


def add_(dst, src, n, acc):
    _LAUNCHES[0] += 1
    _store(_flat(dst)[:n], _flat(src)[:n].float(), acc)


# ------------------------------------------------------------------------------------------------------ CCU
def ccu_stats(xb, u, arg, B, HW, C, ws):
    _LAUNCHES[0] += 2
    v = _flat(xb).view(B, HW, C).float()
    mx, idx = v.max(1)
    u.copy_(torch.stack([mx, v.mean(1), v.std(1, unbiased=False)], -1))
    arg.copy_(idx.to(torch.int32))


def _ccu_mlp(u, fc1, fc2, gamma, beta, eps, stats=None):
    C = u.shape[1]
    z1 = torch.einsum("cjk,bck->bcj", fc1.view(C, 3, 3), u)
    z2 = torch.einsum("cj,bcj->bc", fc2.view(C, 3), F.relu(z1))
    if gamma is not None:
        mu, var = stats if stats is not None else (z2.mean(0), z2.var(0, unbiased=False))
        z2n = (z2 - mu) * torch.rsqrt(var + eps) * gamma + beta
        return torch.sigmoid(z2n), z2, mu, var
    return torch.sigmoid(z2), z2, None, None


def ccu_mlp_fwd(u, fc1, fc2, gamma, beta, rmean, rvar, nbt, momentum, eps, gate, save, B, C, frozen=False):
    _LAUNCHES[0] += 1
    frozen = frozen and gamma is not None
    g, z2, mu, var = _ccu_mlp(u, fc1, fc2, gamma, beta, eps, (rmean.clone(), rvar.clone()) if frozen else None)
    gate.copy_(g)
    if frozen:                                   # (emulation only: the statistics travel to the backward in two free save slots)
        sv = save.view(B, C, 8)
        sv[:, :, 6] = rmean; sv[:, :, 7] = rvar
        return
    if gamma is not None:
        rmean.mul_(1 - momentum).add_(momentum * mu)
        rvar.mul_(1 - momentum).add_(momentum * var * (B / max(B - 1, 1)))
        nbt.add_(1)


def ccu_mlp_bwd(dgate, u, fc1, fc2, gamma, beta, save, du, dfc1, dfc2, dgamma, dbeta, B, C, frozen=False):
    _LAUNCHES[0] += 1
    uu = u.detach().clone().requires_grad_(True)
    f1 = fc1.detach().clone().requires_grad_(True)
    f2 = fc2.detach().clone().requires_grad_(True)
    if gamma is not None:
        ga = gamma.detach().clone().requires_grad_(True)
        be = beta.detach().clone().requires_grad_(True)
        sv = save.view(B, C, 8)
        g = _ccu_mlp(uu, f1, f2, ga, be, 1e-5, (sv[0, :, 6].clone(), sv[0, :, 7].clone()) if frozen else None)[0]
        a, b, c, d, e = torch.autograd.grad(g, (uu, f1, f2, ga, be), dgate)
        dgamma.copy_(d)
        dbeta.copy_(e)
    else:
        g = _ccu_mlp(uu, f1, f2, None, None, 1e-5)[0]
        a, b, c = torch.autograd.grad(g, (uu, f1, f2), dgate)
        dgamma.zero_()
        dbeta.zero_()
    du.copy_(a)
    dfc1.copy_(b)
    dfc2.copy_(c)


def ccu_dgate(dx1, xb, dgate, B, HW, C, ws):
    _LAUNCHES[0] += 2
    dgate.copy_((_flat(dx1).view(B, HW, C).float() * _flat(xb).view(B, HW, C).float()).sum(1))


def ccu_apply_bwd(dx1, xb, gate, u, arg, du, dxb, acc, B, HW, C):
    _LAUNCHES[0] += 1
    g = _flat(dx1).view(B, HW, C).float() * gate.view(B, 1, C)
    x = _flat(xb).view(B, HW, C).float()
    mean, std = u[:, :, 1].view(B, 1, C), u[:, :, 2].view(B, 1, C)
    g = g + du[:, :, 1].view(B, 1, C) / HW + du[:, :, 2].view(B, 1, C) * (x - mean) / (HW * std)
    onehot = torch.zeros(B, HW, C)
    onehot.scatter_(1, arg.long().view(B, 1, C), 1.0)
    g = g + onehot * du[:, :, 0].view(B, 1, C)
    _store(_flat(dxb).view(B, HW, C), g, acc)


# ------------------------------------------------------------------------------------------------------ SRM
def row_stats_arg(x, u, arg, M, C):
    _LAUNCHES[0] += 1
    xf = x.float()
    mx, idx = xf.max(1)
    u.copy_(torch.stack([mx, xf.mean(1), xf.std(1, unbiased=True)], 1))
    arg.copy_(idx.to(torch.int32))


def _srm(u, pw, dw, gamma, beta, eps, B, H, W, stats=None):
    uf = u.view(B, H, W, 3).permute(0, 3, 1, 2)
    f = F.gelu(F.conv2d(uf, pw.view(1, 3, 1, 1)) + F.conv2d(uf, dw.view(1, 3, 3, 3), padding=1))
    mu, var = f.mean(), f.var(unbiased=False)
    if stats is not None:                        # frozen: (mean, rstd) constants
        fn = (f - stats[0]) * stats[1] * gamma + beta
        return torch.sigmoid(fn).reshape(-1), mu, var
    fn = (f - mu) * torch.rsqrt(var + eps) * gamma + beta
    return torch.sigmoid(fn).reshape(-1), mu, var


def srm_fwd(u, pw, dw, gamma, beta, rmean, rvar, nbt, momentum, eps, gm, save, st, B, H, W, ws, frozen=False):
    _LAUNCHES[0] += 3
    if frozen:
        st.view(-1)[0] = rmean.view(-1)[0]
        st.view(-1)[1] = torch.rsqrt(rvar.view(-1)[0] + eps)
        gm.copy_(_srm(u, pw, dw, gamma, beta, eps, B, H, W, (st.view(-1)[0].clone(), st.view(-1)[1].clone()))[0])
        return
    g, mu, var = _srm(u, pw, dw, gamma, beta, eps, B, H, W)
    gm.copy_(g)
    n = B * H * W
    rmean.mul_(1 - momentum).add_(momentum * mu)
    rvar.mul_(1 - momentum).add_(momentum * var * (n / max(n - 1, 1)))
    nbt.add_(1)


def row_dot(a, b, out, M, C):
    _LAUNCHES[0] += 1
    out.copy_((a.float() * b.float()).sum(1))


def srm_bwd(dgm, u, gm, save, st, pw, dw, gamma, beta, du, dpw, ddw, dgamma, dbeta, B, H, W, ws, frozen=False):
    _LAUNCHES[0] += 4
    uu = u.detach().clone().requires_grad_(True)
    p = pw.detach().clone().requires_grad_(True)
    d = dw.detach().clone().requires_grad_(True)
    ga = gamma.detach().clone().requires_grad_(True)
    be = beta.detach().clone().requires_grad_(True)
    g = _srm(uu, p, d, ga, be, 1e-5, B, H, W, (st.view(-1)[0].clone(), st.view(-1)[1].clone()) if frozen else None)[0]
    a, b, c, e, f = torch.autograd.grad(g, (uu, p, d, ga, be), dgm)
    du.copy_(a)
    dpw.copy_(b)
    ddw.copy_(c)
    dgamma.copy_(e)
    dbeta.copy_(f)


def srm_apply_bwd(dh3, h2, z, gm, u, arg, du, dz, M, C):
    _LAUNCHES[0] += 1
    x = h2.float()
    g = dh3.float() * gm.view(M, 1)
    mean, std = u[:, 1].view(M, 1), u[:, 2].view(M, 1)
    g = g + du[:, 1].view(M, 1) / C + du[:, 2].view(M, 1) * (x - mean) / ((C - 1) * std)
    onehot = torch.zeros(M, C)
    onehot.scatter_(1, arg.long().view(M, 1), 1.0)
    g = g + onehot * du[:, 0].view(M, 1)
    zf = z.float().detach().requires_grad_(True)
    dz.copy_(torch.autograd.grad(F.gelu(zf), zf, g)[0])


# ------------------------------------------------------------------------------------------------------ elementwise
def silu_mul_fwd(g, v, out, n):
    _LAUNCHES[0] += 1
    out.copy_(F.silu(g.float()) * F.silu(v.float()))


def silu_mul_bwd(dout, g, v, dg, dv, n):
    _LAUNCHES[0] += 1
    gf = g.float().detach().requires_grad_(True)
    vf = v.float().detach().requires_grad_(True)
    a, b = torch.autograd.grad(F.silu(gf) * F.silu(vf), (gf, vf), dout.float())
    dg.copy_(a)
    dv.copy_(b)


def _ls(x, y, p, s, t, ls, w):
    pz = p * s + t if s is not None else p
    inner = (1 - w) * y + w * pz if y is not None else pz
    return x + ls.view(1, -1) * inner


def ls_combine_fwd(x, y, p, s, t, ls, w, out, M, C):
    _LAUNCHES[0] += 1
    out.copy_(_ls(x.float(), None if y is None else y.float(), p.float(), s, t, ls, w))


def ls_combine_bwd(dout, y, p, s, t, ls, w, dy, acc_dy, dp, dls, dw, M, C, ws):
    _LAUNCHES[0] += 2
    pz = (p.float() * s + t) if s is not None else p.float()
    pz = pz.detach().requires_grad_(True)
    l = ls.detach().clone().requires_grad_(True)
    if y is not None:
        yf = y.float().detach().requires_grad_(True)
        ww = w.detach().clone().requires_grad_(True)
        o = l.view(1, -1) * ((1 - ww) * yf + ww * pz)
        a, b, c, d = torch.autograd.grad(o, (yf, pz, l, ww), dout.float())
        _store(dy, a, acc_dy)
        dw.copy_(d)
    else:
        o = l.view(1, -1) * pz
        b, c = torch.autograd.grad(o, (pz, l), dout.float())
    dp.copy_(b)
    dls.copy_(c.view(dls.shape))


# ------------------------------------------------------------------------------------------------------ resampling
def gather_cast(src, idx, dst):
    i = idx.long()
    dst.copy_(torch.where(i > 0, src[(i - 1).clamp_min(0)], torch.zeros((), dtype=src.dtype)).to(dst.dtype))


def make_tables(Mh, Mw, dev):
    """test emulation keeps the dense per-axis matrices [Ho,Hi], [Wo,Wi]"""
    return dict(Mh=Mh.float().contiguous(), Mw=Mw.float().contiguous())


def resample(x, y, B, Hi, Wi, Ho, Wo, C, tables, ldx=None, x_off=0, ldy=None, y_off=0, acc=False):
    _LAUNCHES[0] += 1
    ldx = C if ldx is None else ldx
    ldy = C if ldy is None else ldy
    xi = _as(_flat(x), (B, Hi, Wi, C), (Hi * Wi * ldx, Wi * ldx, ldx, 1), x_off).float()
    o = torch.einsum("ih,bhwc,jw->bijc", tables["Mh"], xi, tables["Mw"])
    _store(_as(_flat(y), (B, Ho, Wo, C), (Ho * Wo * ldy, Wo * ldy, ldy, 1), y_off), o, acc)


def maxpool2_scale_bwd(dz, lddz, coff, rb, w, drb, dw, B, H, W, C, ws):
    _LAUNCHES[0] += 2
    g = _as(_flat(dz), (B, H // 2, W // 2, C), ((H // 2) * (W // 2) * lddz, (W // 2) * lddz, lddz, 1), coff).float()
    x = _flat(rb).view(B, H, W, C).float().permute(0, 3, 1, 2).detach().requires_grad_(True)
    ww = w.detach().clone().requires_grad_(True)
    o = F.max_pool2d(x, 2) * ww.view(1, C, 1, 1)
    a, b = torch.autograd.grad(o, (x, ww), g.permute(0, 3, 1, 2))
    _flat(drb).view(B, H, W, C).copy_(a.permute(0, 2, 3, 1))
    dw.copy_(b.view(dw.shape))


def head_upsample_bwd(dlogits, dyh, B, h, w, ncls):
    _LAUNCHES[0] += 1
    y = torch.zeros(B, ncls, h, w, requires_grad=True)
    o = F.interpolate(y, scale_factor=2, mode="bilinear")
    g = torch.autograd.grad(o, y, dlogits.float())[0]
    dyh.view(B, h, w, ncls).copy_(g.permute(0, 2, 3, 1))


def adamw(p, g, m, v, n, hyper):
    _LAUNCHES[0] += 1
    lr, b1, b2, eps, wd, step = [float(h) for h in hyper[:6]]
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)
