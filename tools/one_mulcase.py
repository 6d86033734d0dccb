"""Epilogue variants of the narrow decoder GEMM (M=200704, N=64, K=64) timed with CUDA events, L2 flushed.
B200: plain 27.3, SiLU 27.0, mul 29.3, mul + SiLU 37.5, SiLU(g)*SiLU(v) 44.4, residual 27.6 us (roofline 8-12 us): these launches are
bound by the 4 epilogue warps per CTA (168 registers per thread allow 2 CTAs x 192 threads per SM, not 8 epilogue warps each).
Tried: 96-register "lean" instantiations with 2 x (2 + 8) warps per SM -- this shape 26.7 -> 23.4 us (SiLU*SiLU 44.4 -> 33.5), but the
register cap slows the 4-warp launches by as much: forward 10.25 ms either way; everywhere (convs included) 10.43 ms.  Not kept."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
M, N, K = 200704, 64, 64
DEV = "cuda:0"
a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(N, device=DEV)
out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
r = torch.randn(M, N, device=DEV).to(torch.bfloat16)
flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(x.elapsed_time(y) for x, y in ev) / n * 1e3
for name, kw in (("plain", {}), ("silu", dict(act=ops.ACT_SILU)), ("mul", dict(mul=r, ldmul=N)), ("mul+silu act", dict(act=ops.ACT_SILU, mul=r, ldmul=N)),
                 ("silu*silu", dict(act=ops.ACT_SILU, mul=r, ldmul=N, mul_act=ops.ACT_SILU)), ("res1", dict(res1=r, ldr1=N))):
    print(name, round(timeit(lambda: ops.linear(a, w, out, bias=bias, impl=ops.GEMM_TCGEN05, **kw)), 1), "us")
