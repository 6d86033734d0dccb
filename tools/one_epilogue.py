"""Which epilogue option of gemm_tc is slow?  python tools/one_epilogue.py   (M=200704, N=64, K=64 and M=3136, N=512, K=512)"""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
def timeit(fn, n=5, inner=8):
    """median over n measurements of `inner` back-to-back launches (the queue hides the host launch latency)"""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return sorted(ts)[len(ts) // 2]
for M, N, K in ((200704, 64, 64), (75264, 512, 64), (50176, 128, 128)):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    r2 = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    cs = torch.rand(N, device=DEV)
    rs = torch.rand(M, device=DEV)
    cases = {
        "plain": dict(),
        "res1": dict(res1=r, ldr1=N),
        "res1+cscale": dict(res1=r, ldr1=N, res1_cscale=cs),
        "res1+cscale+res2": dict(res1=r, ldr1=N, res1_cscale=cs, res2=r2, ldr2=N),
        "act=relu": dict(act=ops.ACT_RELU),
        "act=silu": dict(act=ops.ACT_SILU),
        "act=gelu": dict(act=ops.ACT_GELU),
        "silu*silu(mul)": dict(act=ops.ACT_SILU, mul=r, ldmul=N, mul_act=ops.ACT_SILU),
        "mul only": dict(mul=r, ldmul=N),
        "mul gelu_grad": dict(mul=r, ldmul=N, mul_act=6),
        "row_scale+res1": dict(row_scale=rs, res1=r, ldr1=N),
    }
    for name, kw in cases.items():
        ms = timeit(lambda: ops.linear(a, w, out, bias=bias, impl=ops.GEMM_TCGEN05, **kw))
        nb = M * K * 2 + N * K * 2 + M * N * 2 * (1 + sum(k in kw for k in ("res1", "res2", "mul")))
        print(f"M={M} N={N} K={K} {name:20s} {ms*1e3:8.1f} us  {nb / ms / 1e9:6.2f} TB/s")
