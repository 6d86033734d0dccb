"""Micro-benchmark of the GEMM / conv shapes of the Synapse B=64 forward (device time, CUDA events)."""
import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
DEV = "cuda:0"
B = 64
shapes = [  # (name, M, N, K, residual)
    ("s1.q/proj", B * 3136, 64, 64, True), ("s1.fc1", B * 3136, 512, 64, False), ("s1.fc2", B * 3136, 64, 512, True),
    ("s2.q", B * 784, 128, 128, True), ("s2.fc1", B * 784, 1024, 128, False), ("s2.fc2", B * 784, 128, 1024, True),
    ("s3.q", B * 196, 320, 320, True), ("s3.fc1", B * 196, 1280, 320, False), ("s3.fc2", B * 196, 320, 1280, True),
    ("s4.q", B * 49, 512, 512, True), ("s4.fc1", B * 49, 2048, 512, False), ("s4.fc2", B * 49, 512, 2048, True),
    ("s1.sr", B * 49, 64, 4096, False), ("se1.qkv", B * 3136, 384, 128, False), ("se1.mixer", B * 3136, 64, 128, True),
    ("dec1.fc1", B * 3136, 256, 64, False), ("dec1.tpg", B * 3136, 192, 64, False), ("head.logits", B * 12544, 9, 64, False),
]
flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
print(f"{'name':12s} {'M':>8s} {'N':>5s} {'K':>5s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
for name, M, N, K, res in shapes:
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    odt = torch.float32 if name == "head.logits" else torch.bfloat16
    out = torch.empty(M, N, device=DEV, dtype=odt)
    r = torch.randn(M, N, device=DEV).to(torch.bfloat16) if res else None
    fn = lambda: ops.linear(a, w, out, bias=bias, res1=r, ldr1=N if res else 0, impl=ops.GEMM_TCGEN05)
    ms = timeit(fn)
    byt = M * K * 2 + N * K * 2 + M * N * out.element_size() * (1) + (M * N * 2 if res else 0)
    print(f"{name:12s} {M:8d} {N:5d} {K:5d} {ms:8.3f} {2*M*N*K/ms/1e9:8.1f} {byt/ms/1e6:8.0f}")
for name, Cin, Cout, k, H in (("rb.conv2 5x5", 32, 32, 5, 224), ("out.conv 3x3", 64, 64, 3, 112), ("up.conv 3x3", 64, 32, 3, 112)):
    x = torch.randn(B, H, H, Cin, device=DEV).to(torch.bfloat16)
    wm = (torch.randn(Cout, k * k * Cin, device=DEV) / 30).to(torch.bfloat16)
    bias = torch.randn(Cout, device=DEV)
    out = torch.empty(B, H, H, Cout, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.conv_nhwc(x, wm, out, k, 1, k // 2, bias=bias, act=ops.ACT_LEAKY, slope=0.01, impl=ops.GEMM_TCGEN05)
    ms = timeit(fn)
    fl = 2.0 * B * H * H * Cout * Cin * k * k
    byt = x.numel() * 2 + out.numel() * 2
    print(f"{name:12s} {B*H*H:8d} {Cout:5d} {k*k*Cin:5d} {ms:8.3f} {fl/ms/1e9:8.1f} {byt/ms/1e6:8.0f}")
