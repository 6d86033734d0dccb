"""Strided convs of the encoder at a given batch: implicit (element-strided TMA boxes) vs im2col + GEMM, error and time.
usage: python tools/one_strided.py B"""
import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
B = int(sys.argv[1])
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        flush.zero_(); e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n * 1e3
ws = torch.zeros(1 << 24, device="cuda")
for (Cin, Cout, k, st, pd, H) in ((64, 64, 8, 8, 0, 56), (128, 128, 4, 4, 0, 28), (320, 320, 2, 2, 0, 14), (64, 128, 3, 2, 1, 56),
                                  (128, 320, 3, 2, 1, 28), (320, 512, 3, 2, 1, 14)):
    W = H
    Ho = (H + 2 * pd - k) // st + 1
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    wm = (torch.randn(Cout, k * k * Cin, device="cuda") / math.sqrt(k * k * Cin)).to(torch.bfloat16)
    bias = torch.randn(Cout, device="cuda")
    out = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty(B * Ho * Ho, Cout, device="cuda", dtype=torch.bfloat16)
    col = torch.empty(B * Ho * Ho, k * k * Cin, device="cuda", dtype=torch.bfloat16)
    def implicit(w=ws):
        ops.conv_nhwc(x, wm, out, k, st, pd, bias=bias, impl=ops.GEMM_TCGEN05, split_ws=w)
    def explicit():
        ops.im2col(x, col, B, H, W, Cin, k, st, pd, Ho, Ho, k * k * Cin)
        ops.gemm(col, wm, out2, M=B * Ho * Ho, N=Cout, K=k * k * Cin, lda=k * k * Cin, ldw=k * k * Cin, ldc=Cout, bias=bias,
                 impl=ops.GEMM_TCGEN05, split_ws=ws)
    explicit(); implicit(); torch.cuda.synchronize()
    e1 = ((out.reshape(-1, Cout).float() - out2.float()).norm() / out2.float().norm()).item()
    implicit(None); torch.cuda.synchronize()
    e2 = ((out.reshape(-1, Cout).float() - out2.float()).norm() / out2.float().norm()).item()
    print(f"B{B} Cin{Cin} Cout{Cout} k{k}s{st} {H}x{W}: err split {e1:.2e} nosplit {e2:.2e} | implicit {timeit(implicit):.1f} us, "
          f"no split {timeit(lambda: implicit(None)):.1f} us, im2col+gemm {timeit(explicit):.1f} us")
