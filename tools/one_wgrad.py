"""Times the deferred tcgen05 weight-gradient GEMM on the shapes of one ACDC batch-24 training step.
usage: python tools/one_wgrad.py [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import train_ops as tops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = 24
shapes = []   # (tag, M, N, K, rs_div)
for st, (C, hw, sr, mr) in enumerate([(64, 56, 8, 8), (128, 28, 4, 8), (320, 14, 2, 4), (512, 7, 1, 4)], 1):
    M = B * hw * hw
    Mk = B * (hw // sr) ** 2
    shapes += [(f"s{st}.q", M, C, C, 0), (f"s{st}.kv", Mk, 2 * C, C, 0), (f"s{st}.proj", M, C, C, hw * hw),
               (f"s{st}.fc1", M, mr * C, C, 0), (f"s{st}.fc2", M, C, mr * C, hw * hw)]
    if sr > 1:
        shapes.append((f"s{st}.sr", Mk, C, sr * sr * C, 0))
shapes += [("dec1.pw", B * 3136, 64, 64, 0), ("dec1.fc1", B * 3136, 256, 64, 0), ("head.1x1", B * 224 * 224, 32, 32, 0),
           ("stem.k25", B * 224 * 224, 32, 25, 0)]
dev = "cuda:0"
ws = torch.zeros(1 << 27, device=dev)
flush = torch.zeros(64 << 20, device=dev)
tot = 0.0
def graph_time(fn, n):
    """device time of one call: n calls captured in a CUDA graph (no host latency between them), replayed 5 times"""
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(n):
                fn(i)
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)


for tag, M, N, K, rsd in shapes:
    ldx = (K + 7) // 8 * 8
    R = max(2, min(8, int(300e6 // (M * (N + ldx) * 2)) + 1))           # rotate operand copies: > L2 in total where it matters
    dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(R)]
    xs = [torch.randn(M, ldx, device=dev).bfloat16() for _ in range(R)]
    dw, db = torch.zeros(N * K, device=dev), torch.zeros(N, device=dev)
    rs = None
    if rsd:
        rs = ((torch.rand(M // rsd, device=dev) < 0.9).float() / 0.9)
    kw = dict(M=M, N=N, K=K, ldy=N, y_off=0, ldx=ldx, x_off=0, row_scale=rs, rs_div=rsd or 1, rs_binary=rsd > 0, dbias=db, ws=ws)
    jobs, used = tops.gemm_wgrad_partial(dys[0], xs[0], dw, **kw)
    S = jobs[0][3] if jobs else 1
    t_r = 0.0
    if jobs:
        tab, nj, nb = tops.wgrad_reduce_table(jobs)
        tab = tab.to(dev)
        t_r = graph_time(lambda i: tops.wgrad_reduce_batch(tab, nj, nb), 8)
    torch.cuda.synchronize()
    t_p = graph_time(lambda i: tops.gemm_wgrad_partial(dys[i % R], xs[i % R], dw, **kw), 2 * R)
    gf = 2.0 * M * N * K / 1e9
    mb = (M * (N + K) * 2 + N * K * 4) / 1e6
    tot += t_p + t_r
    print(f"{tag:10s} M={M:7d} N={N:5d} K={K:5d} rs={rsd:5d}  S={S:4d}  partial {t_p:7.1f} us  reduce {t_r:6.1f} us   "
          f"{gf / t_p * 1e3:7.1f} TFLOP/s {mb / t_p * 1e3 / 1e3:6.2f} TB/s")
print(f"sum {tot:.1f} us")
