"""Depthwise 3x3 forward (+GELU, pre-activation kept) and filter gradient on one Mix-FFN shape, timed with CUDA events.
usage: python tools/one_dwconv.py B H W C   (CENET_B200_DW_STAGED=0/1 selects the register / shared-memory staged kernels)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops, train_ops as tops
B, H, W, C = (int(a) for a in sys.argv[1:5])
x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
dz = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
y, z = torch.empty_like(x), torch.empty_like(x)
w9 = torch.randn(9, C, device="cuda") / 3
bias = torch.randn(C, device="cuda")
dw, db = torch.zeros(C, 1, 3, 3, device="cuda"), torch.zeros(C, device="cuda")
ws = torch.zeros(1 << 24, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
nbytes = x.numel() * 2
t1 = timeit(lambda: ops.dwconv3x3(x, y, w9, B, H, W, C, bias=bias, act=ops.ACT_GELU))
t2 = timeit(lambda: ops.dwconv3x3(x, y, w9, B, H, W, C, bias=bias, act=ops.ACT_GELU, zout=z))
t3 = timeit(lambda: tops.dwconv3x3_wgrad(x, dz, dw, db, B, H, W, C, 1, False, C, 0, C, 0, ws))
print(f"staged={os.environ.get('CENET_B200_DW_STAGED', '1')} B{B} {H}x{W} C{C}: fwd {t1:.1f} us = {2 * nbytes / t1 / 1e6:.2f} TB/s | "
      f"fwd+z {t2:.1f} us = {3 * nbytes / t2 / 1e6:.2f} TB/s | wgrad {t3:.1f} us = {2 * nbytes / t3 / 1e6:.2f} TB/s (L2 flushed)")
