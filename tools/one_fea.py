"""FEA combine forward (+ backward) on one DSEB level, timed with CUDA events (L2 flushed).
usage: python tools/one_fea.py B C2 H W s0 s1 [s2]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
B, C2, H, W = (int(a) for a in sys.argv[1:5])
scales = [float(a) for a in sys.argv[5:]]
y = torch.randn(B, C2, H, W, device="cuda").to(torch.bfloat16)
g = torch.randn(B, C2, H, W, device="cuda").to(torch.bfloat16)
z = torch.empty_like(y)
w = torch.rand(C2, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
t = timeit(lambda: ops.fea_combine(y, g, z, w, B, C2, H, W, scales))
print(f"fea_combine B{B} C{C2} {H}x{W} scales {scales}: {t:.1f} us = {3 * y.numel() * 2 / t / 1e6:.2f} TB/s algorithmic")
