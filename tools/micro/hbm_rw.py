"""HBM bandwidth by access mix on this GPU (torch kernels; measurement only): read-only, write-only, copy.
Write-heavy kernels (GEMM with K << N, dwconv with a stored pre-activation) should be judged against the mix they have."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for mb in (77, 256, 2048):
    n = mb * (1 << 20) // 2
    a = torch.empty(n, device="cuda", dtype=torch.bfloat16).normal_()
    b = torch.empty_like(a)
    rd = t(lambda: a.view(torch.int16).max())
    wr = t(lambda: b.zero_())
    cp = t(lambda: b.copy_(a))
    print(f"{mb:5d} MB: read-only {mb * 1.048576e6 / rd / 1e12:.2f} TB/s | write-only {mb * 1.048576e6 / wr / 1e12:.2f} TB/s | "
          f"copy {2 * mb * 1.048576e6 / cp / 1e12:.2f} TB/s (read+write)")
