"""One gemm_tc launch with a chosen epilogue for ncu: python tools/one_epi_case.py [relu|mul|plain|res1] [M N K]"""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
case = sys.argv[1] if len(sys.argv) > 1 else "relu"
M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (75264, 512, 64)
DEV = "cuda:0"
a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(N, device=DEV)
out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
r = torch.randn(M, N, device=DEV).to(torch.bfloat16)
kw = {"plain": {}, "relu": dict(act=ops.ACT_RELU), "mul": dict(mul=r, ldmul=N, mul_act=6), "res1": dict(res1=r, ldr1=N)}[case]
for _ in range(3):
    ops.linear(a, w, out, bias=bias, impl=ops.GEMM_TCGEN05, **kw)
torch.cuda.synchronize()
