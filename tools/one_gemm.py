import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.linear(a, w, out, bias=bias, impl=ops.GEMM_TCGEN05)
torch.cuda.synchronize()
