"""Per-kernel parity of the TRAINING entry points on the B200: every C-ABI backward / train-forward kernel against its torch
emulation (tests/fake_train_ops.py, which differentiates the oracle's forward formulas with autograd) on the same seeded
inputs, in fp32 (tolerance 1e-4 .. 1e-3: logic) and bf16 (1e-2: rounding).  Integer outputs are compared bit-exactly."""
import math

import pytest
import torch

import fake_ops
import fake_train_ops as FT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F32, BF16 = torch.float32, torch.bfloat16
TOL = {F32: 2e-4, BF16: 1.5e-2}


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def rn(shape, dtype=F32, seed=0, scale=1.0):
    return (torch.randn(shape, generator=gen(seed)) * scale).to(dtype)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _to(v, dev):
    if isinstance(v, torch.Tensor):
        return v.clone().to(dev)
    if isinstance(v, dict):
        return {k: _to(x, dev) for k, x in v.items()}
    return v


def run_pair(mod_fake, mod_real, name, args, kwargs):
    ca, ck = [_to(a, "cpu") for a in args], {k: _to(v, "cpu") for k, v in kwargs.items()}
    ga, gk = [_to(a, DEV) for a in args], {k: _to(v, DEV) for k, v in kwargs.items()}
    getattr(mod_fake, name)(*ca, **ck)
    getattr(mod_real, name)(*ga, **gk)
    torch.cuda.synchronize()
    return ca, ck, ga, gk


def check(name, args, kwargs, outs, tol, real=None, fake=None):
    from cenet_b200 import train_ops as tops
    ca, ck, ga, gk = run_pair(fake or FT, real or tops, name, args, kwargs)
    for o in outs:
        c = ck[o] if isinstance(o, str) else ca[o]
        g = gk[o] if isinstance(o, str) else ga[o]
        if c.dtype in (torch.int32, torch.int64):
            assert torch.equal(c, g.cpu()), (name, o)
        else:
            e = rel(g, c)
            assert e < tol, (name, o, e)


def ws():
    return torch.zeros(1 << 22, dtype=F32)


# ------------------------------------------------------------------------------------------------------ GEMM extensions
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_gemm_training_epilogue(dtype):
    from cenet_b200 import ops
    M, N, K = 300, 72, 64
    a, w, out = rn((M, K), dtype, 1), rn((N, K), dtype, 2, 0.2), torch.zeros(M, N, dtype=dtype)
    res, mul = rn((M, N), dtype, 3), rn((M, N), dtype, 4)
    rs = torch.rand(3, generator=gen(5)) + 0.5
    kw = dict(M=M, N=N, K=K, lda=K, ldw=K, ldc=N, bias=rn((N,), F32, 6), post_rs=rs, post_rs_div=100, res1=res, ldr1=N,
              mul=mul, ldmul=N, mul_act=fake_ops.ACT_GELU_GRAD, row_scale=rs, rs_div=100)
    check("gemm", [a, w, out], kw, [2], TOL[dtype], real=ops, fake=fake_ops)


def test_gemm_a_mmajor_accumulate():
    from cenet_b200 import ops
    M, N, K = 70, 40, 130                                   # A stored [K, M]
    a, w, out = rn((K, M), F32, 1), rn((K, N), F32, 2), rn((M, N), F32, 3)
    kw = dict(M=M, N=N, K=K, lda=M, a_mmajor=True, ldw=N, w_nmajor=True, ldc=N, res1=out, ldr1=N, impl=fake_ops.GEMM_SIMT)
    check("gemm", [a, w, out], kw, [2], 2e-5, real=ops, fake=fake_ops)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_dwconv_train_zout(dtype):
    from cenet_b200 import ops
    B, H, W, C = 2, 12, 10, 64
    x, out, z = rn((B, H, W, C), dtype, 1), torch.zeros(B, H, W, C, dtype=dtype), torch.zeros(B, H, W, C, dtype=dtype)
    check("dwconv3x3", [x, out, rn((9, C), F32, 2, 0.3), B, H, W, C], dict(bias=rn((C,), F32, 3), act=fake_ops.ACT_GELU, zout=z),
          [1, "zout"], TOL[dtype], real=ops, fake=fake_ops)


# ------------------------------------------------------------------------------------------------------ wgrad
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("M,N,K,T", [(3136, 64, 512, 1), (1000, 512, 64, 1), (777, 20, 24, 1), (2048, 32, 800, 25), (640, 9, 64, 1)])
def test_gemm_wgrad(dtype, M, N, K, T):
    dy, x = rn((M, N + 8), dtype, 1), rn((M, K + 8), dtype, 2)
    dw, db = torch.zeros(N * K), torch.zeros(N)
    rs = torch.rand(M // 64 + 1, generator=gen(3))
    kw = dict(M=M, N=N, K=K, ldy=N + 8, y_off=4, ldx=K + 8, x_off=8, T=T, row_scale=rs, rs_div=64, dbias=db, ws=ws())
    check("gemm_wgrad", [dy, x, dw], kw, [2, "dbias"], 3e-4 if dtype == F32 else 1.5e-2)


@pytest.mark.parametrize("Cin,N,k", [(32, 32, 5), (64, 64, 3), (64, 32, 3)])
def test_conv_wgrad_implicit_im2col(Cin, N, k):
    B, H, W = 2, 24, 20
    dy, x4 = rn((B * H * W, N), BF16, 1), rn((B, H, W, Cin), BF16, 2)
    check("conv_wgrad", [dy, x4, torch.zeros(N, Cin, k, k), k, ws()], {}, [2], 1.5e-2)


@pytest.mark.parametrize("M,N,K,ldy,ldx,rsdiv", [(3 * 3136, 64, 512, 64, 512, 3136), (4 * 784, 1024, 128, 1024, 128, 784),
                                                 (4704, 1280, 320, 1280, 320, 0), (75264, 64, 64, 192, 64, 0), (1176, 512, 2048, 512, 2048, 49 * 4), (1176, 512, 2048, 512, 2048, 49), (1176, 512, 512, 512, 512, 49),
                                                 (5000, 320, 100, 328, 104, 0)])
def test_gemm_wgrad_tcgen05(M, N, K, ldy, ldx, rsdiv):
    """shapes that take the tcgen05 MN-major kernel (bf16, aligned), with and without a per-sample row scale"""
    dy, x = rn((M, ldy), BF16, 1), rn((M, ldx), BF16, 2)
    dw, db = torch.zeros(N * K), torch.zeros(N)
    rs = (torch.rand(M // rsdiv, generator=gen(3)) + 0.5) if rsdiv else None
    kw = dict(M=M, N=N, K=K, ldy=ldy, y_off=0, ldx=ldx, x_off=0, row_scale=rs, rs_div=rsdiv or 1, dbias=db, ws=torch.zeros(1 << 24))
    check("gemm_wgrad", [dy, x, dw], kw, [2, "dbias"], 1.5e-2)


def test_gemm_wgrad_deferred_batch():
    """several weight-gradient GEMMs leave their split partials in one workspace; ONE cenet_wgrad_reduce_batch launch reduces
    them all.  Covers: DropPath mask mode (binary row scale, dropped samples skipped), bias through the ones-MMA, a direct
    single-split result, the tap permutation (T=25) of an im2col'ed conv with K padded to the pitch, the mma.sync path
    (K=3), and a general per-sample scale."""
    from cenet_b200 import train_ops as tops
    keep = 0.8
    cases = [  # M, N, K, ldx, T, rs_div, binary, bias
        (6 * 3136, 64, 256, 256, 1, 3136, True, True),
        (6 * 196, 1280, 320, 320, 1, 196, True, True),
        (24 * 49, 512, 2048, 2048, 1, 49, True, True),
        (4704, 320, 320, 320, 1, 0, False, True),
        (640, 96, 64, 64, 1, 0, False, True),                 # few chunks: may run as one split (direct)
        (20000, 32, 25, 32, 25, 0, False, False),             # stem 5x5: K = 25 taps x 1 channel in a 32-wide im2col buffer
        (9000, 32, 3, 3, 1, 0, False, False),                 # mma.sync fallback (pitch not a multiple of 8)
        (4 * 784, 128, 128, 128, 1, 784, False, True),        # general per-sample scale
    ]
    wsb = torch.zeros(1 << 25, device=DEV)
    off, jobs, outs = 0, [], []
    for i, (M, N, K, ldx, T, rsdiv, binary, bias) in enumerate(cases):
        dy, x = rn((M, N), BF16, 10 + i), rn((M, ldx), BF16, 20 + i)
        rs = None
        if rsdiv:
            rs = torch.rand(M // rsdiv, generator=gen(30 + i))
            rs = ((rs < keep).float() / keep) if binary else rs + 0.5
            if binary:
                rs[1] = 0.0
        dw_c, db_c = torch.zeros(N * K), torch.zeros(N)
        FT.gemm_wgrad(dy, x, dw_c, M=M, N=N, K=K, ldy=N, y_off=0, ldx=ldx, x_off=0, T=T, row_scale=rs, rs_div=rsdiv or 1,
                      dbias=db_c if bias else None)
        dw_g, db_g = torch.zeros(N * K, device=DEV), torch.zeros(N, device=DEV)
        j, used = tops.gemm_wgrad_partial(dy.to(DEV), x.to(DEV), dw_g, M=M, N=N, K=K, ldy=N, y_off=0, ldx=ldx, x_off=0, T=T,
                                          row_scale=None if rs is None else rs.to(DEV), rs_div=rsdiv or 1, rs_binary=binary,
                                          dbias=db_g if bias else None, ws=wsb[off:])
        off += (used + 3) // 4 * 4
        jobs += j
        outs.append((dw_c, db_c if bias else None, dw_g, db_g))
    assert jobs and off < wsb.numel()
    tab, nj, nb = tops.wgrad_reduce_table(jobs)
    tops.wgrad_reduce_batch(tab.to(DEV), nj, nb)
    torch.cuda.synchronize()
    for i, (dw_c, db_c, dw_g, db_g) in enumerate(outs):
        assert rel(dw_g, dw_c) < 1.5e-2, (i, rel(dw_g, dw_c))
        if db_c is not None:
            assert rel(db_g, db_c) < 1.5e-2, (i, "bias", rel(db_g, db_c))


def test_gemm_wgrad_mask_all_dropped():
    """every sample dropped: the accumulator is never written -- the result must be exact zeros, not TMEM garbage"""
    from cenet_b200 import train_ops as tops
    M, N, K = 4 * 196, 320, 320
    dy, x = rn((M, N), BF16, 1).to(DEV), rn((M, K), BF16, 2).to(DEV)
    dw, db = torch.ones(N * K, device=DEV), torch.ones(N, device=DEV)
    wsb = torch.zeros(1 << 22, device=DEV)
    jobs, _ = tops.gemm_wgrad_partial(dy, x, dw, M=M, N=N, K=K, ldy=N, y_off=0, ldx=K, x_off=0, row_scale=torch.zeros(4, device=DEV),
                                      rs_div=196, rs_binary=True, dbias=db, ws=wsb)
    if jobs:
        tab, nj, nb = tops.wgrad_reduce_table(jobs)
        tops.wgrad_reduce_batch(tab.to(DEV), nj, nb)
    torch.cuda.synchronize()
    assert float(dw.abs().max()) == 0.0 and float(db.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_colsum_row_scale_smallk(dtype):
    rows, C = 5000, 32
    x = rn((rows, C), dtype, 1)
    rs = torch.rand(rows, generator=gen(2)) + 0.5
    check("colsum", [x, torch.zeros(C)], dict(rows=rows, C=C, ld=C, row_scale=rs, ws=ws()), [1], 3e-4 if dtype == F32 else 1.5e-2)
    check("colsum", [x, torch.zeros(C)], dict(rows=rows, C=C, ld=C, ws=ws()), [1], 3e-4 if dtype == F32 else 1.5e-2)
    check("row_scale", [x, rs, torch.zeros(rows, C, dtype=dtype), rows, C], {}, [2], TOL[dtype])
    K, N = 4, 64
    dy, w = rn((rows, K), F32, 3), rn((K, N), F32, 4)
    for acc in (False, True):
        check("smallk_dgrad", [dy, w, rn((rows, N), dtype, 5)], dict(rows=rows, K=K, N=N, ldw=N, ldx=N, acc=acc), [2], TOL[dtype])


def test_droppath_mask_kernel():
    """values are 0 or 1/keep, the keep rate is right, a row with keep = 1 never drops, and every launch draws a new mask"""
    from cenet_b200 import train_ops as tops
    n, B = 32, 4096
    keep = torch.linspace(1.0, 0.5, n).to(DEV)
    out = torch.empty(n, B, device=DEV)
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    tops.droppath_mask(out, keep, n, B, 1234, counter)
    a = out.clone()
    tops.droppath_mask(out, keep, n, B, 1234, counter)
    torch.cuda.synchronize()
    assert int(counter.item()) == 2 and not torch.equal(a, out)
    k = keep.view(n, 1)
    assert bool(((a == 0) | ((a - 1 / k).abs() < 1e-6)).all())
    assert bool((a[0] == 1).all())
    rate = (a > 0).float().mean(1)
    assert (rate - keep).abs().max().item() < 0.04                      # 4096 draws per row: sigma <= 0.008
    counter.zero_()
    tops.droppath_mask(out, keep, n, B, 1234, counter)
    assert torch.equal(out, a)                                          # same (seed, counter) -> same mask


def test_gemm_wgrad_mixed_dtypes_unscaled_bias():
    M, N, K = 5000, 4, 64                                   # head: fp32 logits gradient x bf16 activations
    dy, x = rn((M, N), F32, 1), rn((M, K), BF16, 2)
    dw, db = torch.zeros(N * K), torch.zeros(N)
    kw = dict(M=M, N=N, K=K, ldy=N, y_off=0, ldx=K, x_off=0, row_scale=torch.rand(M, generator=gen(3)), dbias=db, bias_unscaled=True,
              ws=ws())
    check("gemm_wgrad", [dy, x, dw], kw, [2, "dbias"], 1.5e-2)


# ------------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("C", [64, 128, 320, 512])
@pytest.mark.parametrize("acc", [False, True])
def test_layernorm_bwd(dtype, C, acc):
    rows = 1000
    dy, x, dx = rn((rows, C), dtype, 1), rn((rows, C), dtype, 2), rn((rows, C), dtype, 3)
    check("layernorm_bwd", [dy, x, rn((C,), F32, 4) + 1, 1e-6, dx, acc, torch.zeros(C), torch.zeros(C), ws()], {}, [4, 6, 7],
          3e-4 if dtype == F32 else 1.5e-2)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("rows,C,ld,off", [(5000, 64, 64, 0), (3000, 20, 72, 24), (100000, 32, 32, 0), (49, 4, 4, 0)])
def test_bn_stats_apply_bwd(dtype, rows, C, ld, off):
    x = (rn((rows, ld), dtype, 1) * 1.5 + 0.3).to(dtype)
    gamma, beta = rn((C,), F32, 2) + 1, rn((C,), F32, 3)
    rm, rv, nbt = rn((C,), F32, 4), torch.rand(C, generator=gen(5)) + 0.5, torch.tensor(7)
    sc, sh, mean, rstd = (torch.zeros(C) for _ in range(4))
    args = [x, rows, C, gamma, beta, rm, rv, nbt, 0.1, 1e-5, sc, sh, mean, rstd, ws()]
    check("bn_stats", args, dict(ldx=ld, x_off=off), [5, 6, 7, 10, 11, 12, 13], 2e-4 if dtype == F32 else 2e-3)
    # apply + backward with the exact statistics
    xv = x.view(-1)[off:].as_strided((rows, C), (ld, 1)).float()
    mean, var = xv.mean(0), xv.var(0, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    sc, sh = gamma * rstd, beta - mean * gamma * rstd
    y = torch.zeros(rows, ld, dtype=dtype)
    for act in (fake_ops.ACT_NONE, fake_ops.ACT_LEAKY):
        check("affine_act", [x, y, rows, C], dict(sa=sc, ta=sh, act=act, slope=0.2, lda=ld, a_off=off, ldo=ld, o_off=off), [1], TOL[dtype])
        FT.affine_act(x, y, rows, C, sa=sc, ta=sh, act=act, slope=0.2, lda=ld, a_off=off, ldo=ld, o_off=off)
        dy, da, dres = rn((rows, ld), dtype, 6), rn((rows, ld), dtype, 7), rn((rows, C), dtype, 8)
        kw = dict(act=act, slope=0.2, acc_da=True, dres=dres, acc_dres=True, ldy=ld, y_off=off, lda=ld, a_off=off, lddres=C, dres_off=0)
        check("bn_bwd", [dy, y if act else None, x, mean, rstd, gamma, rows, C, da, torch.zeros(C), torch.zeros(C), ws()], kw,
              [8, 9, 10, "dres"], 5e-4 if dtype == F32 else 2e-2)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_affine_act_two_branches(dtype):
    rows, C = 4000, 32
    a, b, out = rn((rows, C), dtype, 1), rn((rows, C), dtype, 2), torch.zeros(rows, C, dtype=dtype)
    for sb in (None, rn((C,), F32, 5) + 1):
        kw = dict(sa=rn((C,), F32, 3) + 1, ta=rn((C,), F32, 4), b=b, sb=sb, tb=None if sb is None else rn((C,), F32, 6),
                  act=fake_ops.ACT_LEAKY, slope=0.01)
        check("affine_act", [a, out, rows, C], kw, [1], TOL[dtype])


# ------------------------------------------------------------------------------------------------------ depthwise family
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("C,ldx,xo,dil,up2,bias", [(64, 64, 0, 1, False, True), (20, 64, 20, 3, False, False), (128, 128, 0, 1, True, False)])
def test_dwconv_wgrad(dtype, C, ldx, xo, dil, up2, bias):
    B, H, W = 2, 14, 12
    Hi, Wi = (H // 2, W // 2) if up2 else (H, W)
    x, dz = rn((B, Hi, Wi, ldx), dtype, 1), rn((B, H, W, C + 8), dtype, 2)
    dw, db = torch.zeros(C, 1, 3, 3), torch.zeros(C)
    check("dwconv3x3_wgrad", [x, dz, dw, db if bias else None, B, H, W, C, dil, up2, ldx, xo, C + 8, 0, ws()], {},
          [2, 3] if bias else [2], 3e-4 if dtype == F32 else 1.5e-2)


@pytest.mark.parametrize("B,H,W,C,ldx,xo", [(3, 56, 56, 128, 128, 0), (2, 14, 14, 320, 320, 0), (1, 7, 7, 64, 64, 0),
                                            (2, 28, 30, 64, 192, 64), (1, 20, 130, 64, 64, 0)])
def test_dwconv_wgrad_staged_rows(B, H, W, C, ldx, xo):
    """shared-memory staged filter gradient (bf16, dilation 1): several bands, ragged widths, channel slice, wide rows"""
    x, dz = rn((B, H, W, ldx), BF16, 1), rn((B, H, W, C), BF16, 2)
    dw, db = torch.zeros(C, 1, 3, 3), torch.zeros(C)
    check("dwconv3x3_wgrad", [x, dz, dw, db, B, H, W, C, 1, False, ldx, xo, C, 0, ws()], {}, [2, 3], 2e-3)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_sumpool2_col2im(dtype):
    B, Ho, Wo, C = 2, 7, 9, 40
    full, dx = rn((B, 2 * Ho, 2 * Wo, C), dtype, 1), rn((B, Ho, Wo, C), dtype, 2)
    for acc in (False, True):
        check("sumpool2", [full, dx, B, Ho, Wo, C, acc], {}, [1], TOL[dtype])
    for (H, W, Cin, k, s, p) in ((14, 14, 64, 3, 2, 1), (16, 16, 64, 4, 4, 0), (8, 8, 128, 2, 2, 0)):
        Ho2, Wo2 = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        Kp = k * k * Cin + 8
        dcol, dxi = rn((B * Ho2 * Wo2, Kp), dtype, 3), rn((B, H, W, Cin), dtype, 4)
        for acc in (False, True):
            check("col2im", [dcol, dxi, B, H, W, Cin, k, s, p, Ho2, Wo2, Kp, acc], {}, [1], TOL[dtype])


# ------------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("dqk,dv,vdiv,maps,Nq,Nk", [(64, 64, 1, 2, 200, 49), (64, 64, 1, 1, 784, 784), (16, 32, 2, 8, 196, 196),
                                                    (8, 16, 2, 4, 300, 300), (32, 64, 2, 4, 130, 130), (128, 128, 1, 1, 784, 784),
                                                    (80, 160, 2, 8, 196, 196)])
def test_flash_fwd_bwd(dqk, dv, vdiv, maps, Nq, Nk):
    from cenet_b200 import train_ops as tops
    B = 2
    hv = maps // vdiv
    ldq, ldk, ldv, ldo = maps * dqk + 16, maps * dqk + 8, hv * dv + 8, maps * dv
    Q, K, V = rn((B * Nq, ldq), BF16, 1), rn((B * Nk, ldk), BF16, 2), rn((B * Nk, ldv), BF16, 3)
    O, lse = torch.zeros(B * Nq, ldo, dtype=BF16), torch.zeros(B * maps * Nq)
    scale = dqk ** -0.5
    tail = [B, maps, Nq, Nk, dqk, dv, vdiv, scale, ldq, 8, ldk, 8, ldv, 0, ldo, 0]
    ca, ck, ga, gk = run_pair(FT, tops, "flash_fwd", [Q, K, V, O, lse] + tail, {})
    assert rel(ga[3], ca[3]) < 1.5e-2
    assert (ga[4].cpu() / math.log2(math.e) - ca[4]).abs().max().item() < 2e-2          # kernel stores log2-sum-exp
    # backward on the kernel's own forward results
    Og, lseg = ga[3].cpu(), ga[4].cpu()
    dO = rn((B * Nq, ldo), BF16, 4)
    dQ, dK, dV = torch.zeros_like(Q), torch.zeros_like(K), torch.zeros_like(V)
    delta = torch.zeros(B * maps * Nq)
    ca, ck, ga, gk = run_pair(FT, tops, "flash_bwd", [Q, K, V, Og, dO, lseg, delta, dQ, dK, dV] + tail, {})
    for i, n in ((7, "dQ"), (8, "dK"), (9, "dV")):
        assert rel(ga[i], ca[i]) < 2e-2, (n, rel(ga[i], ca[i]))


@pytest.mark.parametrize("hd,heads,N,B,kmax", [(16, 4, 300, 2, True), (8, 8, 196, 3, True), (8, 8, 784, 2, False), (32, 2, 130, 2, True),
                                               (64, 1, 257, 2, True), (16, 4, 3136, 1, True)])
def test_diffattn_fwd_train_tc(hd, heads, N, B, kmax):
    """training forward of the differential attention on the tcgen05 kernel: per-map outputs and log2 LSE against the emulation
    (and against the per-map mma.sync flash forward), then the flash backward on ITS results"""
    from cenet_b200 import train_ops as tops
    E = 2 * heads * hd
    maps = 2 * heads
    qkv = rn((B * N, 3 * E), BF16, 1)
    Om, lse = torch.zeros(B * N, 2 * E, dtype=BF16), torch.zeros(B * maps * N)
    kw = {"kmax_ws": torch.zeros(B * maps)} if kmax else {}
    ca, ck, ga, gk = run_pair(FT, tops, "diffattn_fwd_train", [qkv, Om, lse, B, N, E, heads], kw)
    assert rel(ga[1], ca[1]) < 1.5e-2, rel(ga[1], ca[1])
    assert (ga[2].cpu() / math.log2(math.e) - ca[2]).abs().max().item() < 2e-2
    if hd == 64:
        return                                                       # (64, 128) has no mma.sync flash instantiation to compare with
    # same operands through the per-map flash forward
    q = qkv.to(DEV)
    O2, lse2 = torch.zeros_like(ga[1]), torch.zeros_like(ga[2])
    tail = [B, maps, N, N, hd, 2 * hd, 2, hd ** -0.5, 3 * E, 0, 3 * E, E, 3 * E, 2 * E, 2 * E, 0]
    tops.flash_fwd(q, q, q, O2, lse2, *tail)
    torch.cuda.synchronize()
    assert rel(ga[1], O2) < 1e-2 and (ga[2] - lse2).abs().max().item() < 5e-3
    dO = rn((B * N, 2 * E), BF16, 4)
    dq, dk, dv = torch.zeros_like(qkv), torch.zeros_like(qkv), torch.zeros_like(qkv)
    delta = torch.zeros(B * maps * N)
    ca, ck, ga2, gk = run_pair(FT, tops, "flash_bwd", [qkv, qkv, qkv, ga[1].cpu(), dO, ga[2].cpu(), delta, dq, dk, dv] + tail, {})
    for i, (c0, c1) in ((7, (0, E)), (8, (E, 2 * E)), (9, (2 * E, 3 * E))):
        assert rel(ga2[i][:, c0:c1], ca[i][:, c0:c1]) < 2e-2, (i, rel(ga2[i][:, c0:c1], ca[i][:, c0:c1]))


@pytest.mark.parametrize("maps,Nq,Nk,B", [(1, 3136, 49, 3), (2, 784, 49, 2), (5, 300, 49, 2), (1, 1000, 64, 2)])
def test_flash_bwd_query_split(maps, Nq, Nk, B):
    """short key sets (SR attention, 49 reduced keys): with a workspace the dK / dV kernel splits the QUERIES over CTAs and a
    fixed-order reduction adds the fp32 partials; same results as the unsplit kernel (bf16 rounding), bit-identical run to run"""
    from cenet_b200 import train_ops as tops
    d = 64
    ld = maps * d
    Q, K, V = rn((B * Nq, ld), BF16, 1).to(DEV), rn((B * Nk, 2 * ld), BF16, 2).to(DEV), None
    O, lse = torch.zeros(B * Nq, ld, dtype=BF16, device=DEV), torch.zeros(B * maps * Nq, device=DEV)
    scale = d ** -0.5
    tail = [B, maps, Nq, Nk, d, d, 1, scale, ld, 0, 2 * ld, 0, 2 * ld, ld, ld, 0]        # K | V interleaved like the kv GEMM output
    tops.flash_fwd(Q, K, K, O, lse, *tail)
    dO = rn((B * Nq, ld), BF16, 4).to(DEV)
    delta = torch.zeros(B * maps * Nq, device=DEV)
    outs = []
    for wsb in (None, torch.zeros(1 << 22, device=DEV), torch.zeros(1 << 22, device=DEV)):
        dQ, dKV = torch.zeros_like(Q), torch.zeros_like(K)
        tops.flash_bwd(Q, K, K, O, dO, lse, delta, dQ, dKV, dKV, *tail, ws=wsb)
        torch.cuda.synchronize()
        outs.append((dQ.clone(), dKV.clone()))
    assert rel(outs[1][1], outs[0][1]) < 1e-2 and torch.equal(outs[1][0], outs[0][0])
    assert torch.equal(outs[1][1], outs[2][1])


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_softmax_bwd_and_diff_rmsnorm(dtype):
    rows, n = 500, 196
    P = torch.softmax(rn((rows, n), F32, 1), -1).to(dtype)
    dP = rn((rows, n), dtype, 2)
    check("softmax_bwd_rows_", [P, dP, rows, n], {}, [1], TOL[dtype])
    from cenet_b200 import train_ops as tops
    # segment = 2 * head_dim of the value heads: 16 / 32 (ACDC, Synapse), 40 (Synapse 14x14), 64 / 128 / 320 (skin)
    for M, heads, seg in ((700, 4, 32), (333, 16, 16), (200, 8, 40), (257, 2, 128), (100, 2, 320)):
        Om, lam = rn((M, 2 * heads * seg), dtype, 3), torch.tensor([0.37, 0, 0, 0])
        o = torch.zeros(M, heads * seg, dtype=dtype)
        check("diff_rmsnorm_fwd", [Om, lam, o, M, heads, seg, 1e-5, 0.44], {}, [2], TOL[dtype])
        do, dOm, dlam = rn((M, heads * seg), dtype, 4), torch.zeros_like(Om), torch.zeros(4)
        ca, ck, ga, gk = run_pair(FT, tops, "diff_rmsnorm_bwd", [do, Om, lam, dOm, dlam, M, heads, seg, 1e-5, 0.44, ws()], {})
        assert rel(ga[3], ca[3]) < TOL[dtype], (M, heads, seg)
        assert abs(ga[4][0].item() - ca[4][0].item()) < 2e-3 * max(1.0, abs(ca[4][0].item())) * (1 if dtype == F32 else 20)


def test_lambda_fwd_bwd():
    hd = 80
    v = [rn((hd,), F32, i, 0.1) for i in range(4)]
    check("lambda_fwd", v + [hd, 0.62, torch.zeros(4)], {}, [], 1e-6)
    from cenet_b200 import train_ops as tops
    lam = torch.zeros(4, device=DEV)
    tops.lambda_fwd(*[t.to(DEV) for t in v], hd, 0.62, lam)
    ref = torch.exp((v[0] * v[1]).sum()) - torch.exp((v[2] * v[3]).sum()) + 0.62
    assert abs(lam[0].item() - ref.item()) < 1e-6
    check("lambda_bwd", [torch.tensor([0.7, 0, 0, 0])] + v + [hd] + [torch.zeros(hd) for _ in range(4)], {}, [6, 7, 8, 9], 1e-5)


# ------------------------------------------------------------------------------------------------------ DSEB
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("scales,H,W", [((1.0, 0.5), 14, 14), ((0.8, 0.4), 28, 28), ((1.0, 0.75, 0.5), 56, 56),
                                         ((1.0, 0.75, 0.5), 128, 128)])      # 128x128: the workspace-resident variant
def test_fea_bwd_matches_autograd_of_oracle_fea(dtype, scales, H, W):
    import torch.nn.functional as F
    from oracle import cenet_oracle as O
    import cenet_b200.train as T
    B, E = 2, 6
    eng = T.TrainEngine.__new__(T.TrainEngine)
    eng.dev = torch.device("cpu")
    T.TrainEngine._TABLES.clear()
    mats = eng._fea_mats(H, W, list(scales))
    y, gate, dz = rn((B, E, H, W), dtype, 1), rn((B, E, H, W), dtype, 2), rn((B, E, H, W), dtype, 3)
    w = rn((E,), F32, 4) + 0.5
    dy, dgate, dw = rn((B, E, H, W), dtype, 5), torch.zeros(B, E, H, W, dtype=dtype), torch.zeros(E)
    for acc in (False, True):
        for mask in (0, sum(1 << i for i, sf in enumerate(scales) if sf == 1.0)):        # with / without the identity shortcut
            check("fea_bwd", [y, gate, dz, w, dy, acc, dgate, dw, B, E, H, W, mats, len(scales), ws()], dict(ident_mask=mask),
                  [4, 6, 7], 5e-4 if dtype == F32 else 2e-2)
    if dtype == F32:                                        # the dense-operator formulation IS the oracle's FEA
        yf = y.clone().requires_grad_(True)
        z = O.fea({"m.w": w.view(1, E, 1, 1)}, "m", yf, list(scales)) + yf + gate * yf
        gy = torch.autograd.grad(z, yf, dz)[0]
        d2, g2, w2 = torch.zeros_like(dy), torch.zeros_like(dgate), torch.zeros(E)
        FT.fea_bwd(y, gate, dz, w, d2, False, g2, w2, B, E, H, W, mats, len(scales), None)
        assert rel(d2, gy) < 1e-5


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_layout_and_add(dtype):
    B, HW, C, Ctot = 2, 100, 40, 80
    x, out = rn((B, Ctot, HW), dtype, 1), rn((B, HW, C), dtype, 2)
    for acc in (False, True):
        check("nchw_to_nhwc_slice", [x, out, B, HW, C, Ctot, 40, acc], {}, [1], TOL[dtype])
        check("add_", [rn((1000,), dtype, 3), rn((1000,), dtype, 4), 1000, acc], {}, [0], TOL[dtype])


# ------------------------------------------------------------------------------------------------------ CFAM pieces
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B", [1, 3])
def test_ccu_train(dtype, B):
    HW, C = 196, 64
    xb = rn((B * HW, C), dtype, 1)
    u, arg = torch.zeros(B, C, 3), torch.zeros(B, C, dtype=torch.int32)
    check("ccu_stats", [xb, u, arg, B, HW, C, ws()], {}, [1, 2], 1e-4 if dtype == F32 else 1e-3)
    FT.ccu_stats(xb, u, arg, B, HW, C, None)
    fc1, fc2 = rn((3 * C, 1, 3), F32, 2, 0.5), rn((C, 3, 1), F32, 3, 0.5)
    bn = B > 1
    gamma, beta = (rn((C,), F32, 4) + 1, rn((C,), F32, 5)) if bn else (None, None)
    rm, rv, nbt = torch.zeros(C), torch.ones(C), torch.tensor(0)
    gate, save = torch.zeros(B, C), torch.zeros(B, C, 8)
    from cenet_b200 import train_ops as tops
    ca, ck, ga, gk = run_pair(FT, tops, "ccu_mlp_fwd", [u, fc1, fc2, gamma, beta, rm, rv, nbt, 0.1, 1e-5, gate, save, B, C], {})
    assert rel(ga[10], ca[10]) < 1e-5
    if bn:
        assert rel(ga[5], ca[5]) < 1e-5 and rel(ga[6], ca[6]) < 1e-5 and int(ga[7]) == 1
    save_g, gate_c = ga[11].cpu(), ca[10]
    dx1 = rn((B * HW, C), dtype, 6)
    dgate = torch.zeros(B, C)
    check("ccu_dgate", [dx1, xb, dgate, B, HW, C, ws()], {}, [2], 1e-4 if dtype == F32 else 2e-3)
    FT.ccu_dgate(dx1, xb, dgate, B, HW, C, None)
    du, d1, d2, dg, db = torch.zeros(B, C, 3), torch.zeros_like(fc1), torch.zeros_like(fc2), torch.zeros(C), torch.zeros(C)
    check("ccu_mlp_bwd", [dgate, u, fc1, fc2, gamma, beta, save_g, du, d1, d2, dg, db, B, C], {}, [7, 8, 9, 10, 11], 2e-4)
    FT.ccu_mlp_bwd(dgate, u, fc1, fc2, gamma, beta, save_g, du, d1, d2, dg, db, B, C)
    dxb = rn((B * HW, C), dtype, 7)
    for acc in (False, True):
        check("ccu_apply_bwd", [dx1, xb, gate_c, u, arg, du, dxb, acc, B, HW, C], {}, [6], TOL[dtype])


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_srm_train(dtype):
    B, H, W, C = 2, 14, 14, 256
    M = B * H * W
    h2 = rn((M, C), dtype, 1)
    u, arg = torch.zeros(M, 3), torch.zeros(M, dtype=torch.int32)
    check("row_stats_arg", [h2, u, arg, M, C], {}, [1, 2], 1e-4 if dtype == F32 else 1e-3)
    FT.row_stats_arg(h2, u, arg, M, C)
    pw, dw = rn((1, 3, 1, 1), F32, 2, 0.5), rn((1, 3, 3, 3), F32, 3, 0.3)
    gamma, beta = torch.tensor([1.3]), torch.tensor([-0.2])
    rm, rv, nbt = torch.zeros(1), torch.ones(1), torch.tensor(0)
    gm, save, st = torch.zeros(M), torch.zeros(M, 2), torch.zeros(4)
    from cenet_b200 import train_ops as tops
    ca, ck, ga, gk = run_pair(FT, tops, "srm_fwd", [u, pw, dw, gamma, beta, rm, rv, nbt, 0.1, 1e-5, gm, save, st, B, H, W, ws()], {})
    assert rel(ga[10], ca[10]) < 1e-4 and rel(ga[5], ca[5]) < 1e-4 and rel(ga[6], ca[6]) < 1e-4
    gm_c, save_g, st_g = ca[10], ga[11].cpu(), ga[12].cpu()
    dh3, z = rn((M, C), dtype, 4), rn((M, C), dtype, 5)
    dgm = torch.zeros(M)
    check("row_dot", [dh3, h2, dgm, M, C], {}, [2], 1e-4 if dtype == F32 else 3e-3)
    FT.row_dot(dh3, h2, dgm, M, C)
    du, dpw, ddw, dg, db = torch.zeros(M, 3), torch.zeros_like(pw), torch.zeros_like(dw), torch.zeros(1), torch.zeros(1)
    check("srm_bwd", [dgm, u, gm_c, save_g, st_g, pw, dw, gamma, beta, du, dpw, ddw, dg, db, B, H, W, ws()], {}, [9, 10, 11, 12, 13], 1e-3)
    FT.srm_bwd(dgm, u, gm_c, save_g, st_g, pw, dw, gamma, beta, du, dpw, ddw, dg, db, B, H, W, None)
    dz = torch.zeros(M, C, dtype=dtype)
    check("srm_apply_bwd", [dh3, h2, z, gm_c, u, arg, du, dz, M, C], {}, [7], TOL[dtype])


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_silu_mul_and_ls_combine(dtype):
    M, C = 3000, 64
    g, v, out = rn((M, C), dtype, 1), rn((M, C), dtype, 2), torch.zeros(M, C, dtype=dtype)
    check("silu_mul_fwd", [g, v, out, M * C], {}, [2], TOL[dtype])
    check("silu_mul_bwd", [rn((M, C), dtype, 3), g, v, torch.zeros_like(g), torch.zeros_like(g), M * C], {}, [3, 4], TOL[dtype])
    x, y, p = rn((M, C), dtype, 4), rn((M, C), dtype, 5), rn((M, C), dtype, 6)
    s, t, ls, w = rn((C,), F32, 7) + 1, rn((C,), F32, 8), rn((1, C, 1, 1), F32, 9) + 0.5, torch.tensor(0.4)
    check("ls_combine_fwd", [x, y, p, s, t, ls, w, out, M, C], {}, [7], TOL[dtype])
    check("ls_combine_fwd", [x, None, p, None, None, ls, None, out, M, C], {}, [7], TOL[dtype])
    dout, dy, dp = rn((M, C), dtype, 10), rn((M, C), dtype, 11), torch.zeros(M, C, dtype=dtype)
    tol = 5e-4 if dtype == F32 else 2e-2
    for acc in (False, True):
        check("ls_combine_bwd", [dout, y, p, s, t, ls, w, dy, acc, dp, torch.zeros(1, C, 1, 1), torch.zeros(()), M, C, ws()], {},
              [7, 9, 10, 11], tol)
    check("ls_combine_bwd", [dout, None, p, None, None, ls, None, None, False, dp, torch.zeros(1, C, 1, 1), None, M, C, ws()], {},
          [9, 10], tol)


# ------------------------------------------------------------------------------------------------------ resampling / head
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_resample_tables_match_torch(dtype):
    import torch.nn.functional as F
    from cenet_b200 import train_ops as tops
    import cenet_b200.train as T
    B, H, W, r, Cc = 2, 28, 28, 8, 128
    eng = T.TrainEngine.__new__(T.TrainEngine)
    eng.dev = torch.device(DEV)
    T.TrainEngine._TABLES.clear()
    tb = eng._pool_tables(H, W)
    x = rn((B, H, W, Cc), dtype, 1)
    pooled = torch.zeros(B * 49, r, dtype=dtype, device=DEV)
    tops.resample(x.to(DEV), pooled, B, H, W, 7, 7, r, tb["pool"], ldx=Cc, x_off=120, ldy=r, y_off=0)
    xs = x[..., 120:128].float().permute(0, 3, 1, 2)
    ref = F.adaptive_avg_pool2d(xs, 7).permute(0, 2, 3, 1).reshape(B * 49, r)
    assert rel(pooled, ref) < TOL[dtype]
    up = torch.zeros(B, H, W, Cc, dtype=dtype, device=DEV)
    tops.resample(pooled, up, B, 7, 7, H, W, r, tb["up"], ldx=r, x_off=0, ldy=Cc, y_off=120)
    pr = pooled.float().cpu().view(B, 7, 7, r).permute(0, 3, 1, 2)
    ref = F.interpolate(F.interpolate(pr, scale_factor=7, mode="bilinear", align_corners=True), size=(H, W), mode="bilinear")
    assert rel(up[..., 120:128], ref.permute(0, 2, 3, 1)) < TOL[dtype]
    # adjoint: <U p, g> == <p, U^T g>
    gq = rn((B, H, W, Cc), dtype, 2).to(DEV)
    back = torch.zeros(B * 49, r, dtype=dtype, device=DEV)
    tops.resample(gq, back, B, H, W, 7, 7, r, tb["up_T"], ldx=Cc, x_off=120, ldy=r, y_off=0)
    lhs = (up[..., 120:128].float() * gq[..., 120:128].float()).sum().item()
    rhs = (pooled.float() * back.float()).sum().item()
    assert abs(lhs - rhs) < (1e-4 if dtype == F32 else 3e-2) * max(1.0, abs(lhs))
    # bilinear x2 align_corners adjoint (UpConv)
    t2 = eng._up2_tables(14, 14)
    gu, dx = rn((B, 28, 28, 64), dtype, 3), torch.zeros(B, 14, 14, 64, dtype=dtype, device=DEV)
    tops.resample(gu.to(DEV), dx, B, 28, 28, 14, 14, 64, t2["up_T"], acc=False)
    xi = torch.zeros(B, 64, 14, 14, requires_grad=True)
    ref = torch.autograd.grad(F.interpolate(xi, scale_factor=2, mode="bilinear", align_corners=True), xi, gu.float().permute(0, 3, 1, 2))[0]
    assert rel(dx, ref.permute(0, 2, 3, 1)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_maxpool_scale_bwd(dtype):
    B, H, W, C = 2, 16, 12, 32
    rb, dz = rn((B, H, W, C), dtype, 1), rn((B, H // 2, W // 2, 2 * C), dtype, 2)
    w = rn((C,), F32, 3) + 0.75
    check("maxpool2_scale_bwd", [dz, 2 * C, C, rb, w, torch.zeros(B, H, W, C, dtype=dtype), torch.zeros(1, C, 1, 1), B, H, W, C, ws()], {},
          [5, 6], 2e-4 if dtype == F32 else 1.5e-2)


def test_head_upsample_bwd_and_adamw():
    B, h, w, ncls = 2, 20, 24, 4
    dl = rn((B, ncls, 2 * h, 2 * w), F32, 1)
    check("head_upsample_bwd", [dl, torch.zeros(B * h * w, ncls), B, h, w, ncls], {}, [1], 1e-5)
    n = 4096
    p, g_, m, v = rn((n,), F32, 2), rn((n,), F32, 3, 0.01), rn((n,), F32, 4, 0.01), torch.rand(n, generator=gen(5)) * 1e-4
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-2, 3.0, 0, 0])
    check("adamw", [p, g_, m, v, n, hyper], {}, [0, 2, 3], 1e-5)
    # against torch.optim.AdamW itself
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([q], lr=1e-3, weight_decay=1e-2)
    q.grad = g_.clone()
    opt.step()
    p2, m2, v2 = p.clone(), torch.zeros(n), torch.zeros(n)
    FT.adamw(p2, g_, m2, v2, n, torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-2, 1.0, 0, 0]))
    assert (p2 - q.detach()).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------------------------------ weight re-pack
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_gather_cast_is_an_exact_indexed_copy(dtype):
    from cenet_b200 import train_ops as tops
    n_src, n = 100003, 4 * 50021
    src = rn((n_src,), F32, 1).to(DEV)
    idx = torch.randint(0, n_src + 1, (n,), generator=gen(2)).to(torch.int32)       # 0 = zero padding
    dst = torch.full((n,), 7.0, dtype=dtype, device=DEV)
    tops.gather_cast(src, idx.to(DEV), dst)
    want = torch.where(idx > 0, src.cpu()[(idx.long() - 1).clamp_min(0)], torch.zeros(())).to(dtype)
    assert torch.equal(dst.cpu(), want)                                              # bit-exact (round-to-nearest cast)
