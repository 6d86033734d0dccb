"""Mix-FFN tail (depthwise 3x3 + GELU + fc2 + residual): fused tcgen05 kernel vs the unfused pair, timed with CUDA events.
usage: python tools/one_mixffn.py B H W Ch C"""
import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
B, H, W, Ch, C = (int(a) for a in sys.argv[1:6])
M = B * H * W
h = torch.randn(B, H, W, Ch, device="cuda").to(torch.bfloat16)
h2 = torch.empty(M, Ch, device="cuda", dtype=torch.bfloat16)
t = torch.randn(M, C, device="cuda")
w9 = torch.randn(9, Ch, device="cuda") / 3
bdw = torch.randn(Ch, device="cuda")
w2 = (torch.randn(C, Ch, device="cuda") / math.sqrt(Ch)).to(torch.bfloat16)
b2 = torch.randn(C, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
def unfused():
    ops.dwconv3x3(h, h2, w9, B, H, W, Ch, bias=bdw, act=ops.ACT_GELU)
    ops.linear(h2, w2, t, bias=b2, res1=t, ldr1=C)
t_un = timeit(unfused)
t_dw = timeit(lambda: ops.dwconv3x3(h, h2, w9, B, H, W, Ch, bias=bdw, act=ops.ACT_GELU))
t_fu = timeit(lambda: ops.mixffn_tail(h, t, w9, bdw, w2, b2, B, H, W, Ch, C))
by = M * Ch * 2 + 2 * M * C * 4
print(f"B{B} {H}x{W} Ch{Ch} C{C}: unfused {t_un:.1f} us (dwconv alone {t_dw:.1f}) | fused {t_fu:.1f} us = {by / t_fu / 1e6:.2f} TB/s "
      f"algorithmic, {2.0 * M * Ch * C / t_fu / 1e6:.0f} TFLOP/s (L2 flushed)")
