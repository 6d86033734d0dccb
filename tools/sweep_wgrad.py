"""Sweeps (tile width, splits) of the tcgen05 weight-gradient kernel on the training shapes (CENET_B200_WGRAD_PLAN override)
and prints the best per shape next to what the plan model picks.  usage: python tools/sweep_wgrad.py"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import train_ops as tops, _lib as L
dev = "cuda:0"
B = 24
shapes = [("s1.q", 75264, 64, 64, 0), ("s1.fc1", 75264, 512, 64, 0), ("s1.fc2", 75264, 64, 512, 3136), ("s2.q", 18816, 128, 128, 0),
          ("s2.fc1", 18816, 1024, 128, 0), ("s2.fc2", 18816, 128, 1024, 784), ("s3.q", 4704, 320, 320, 0), ("s3.fc1", 4704, 1280, 320, 0),
          ("s3.fc2", 4704, 320, 1280, 196), ("s3.kv", 1176, 640, 320, 0), ("s1.sr", 1176, 64, 4096, 0), ("s4.q", 1176, 512, 512, 0),
          ("s4.fc1", 1176, 2048, 512, 0), ("s4.fc2", 1176, 512, 2048, 49), ("dec1.fc1", 75264, 256, 64, 0), ("head.1x1", 1204224, 32, 32, 0)]
ws = torch.zeros(1 << 27, device=dev)


def graph_time(fn, n):
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
    for _ in range(3):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * n)


for tag, M, N, K, rsd in shapes:
    R = max(2, min(8, int(300e6 // (M * (N + K) * 2)) + 1))
    dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(R)]
    xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(R)]
    dw, db = torch.zeros(N * K, device=dev), torch.zeros(N, device=dev)
    rs = ((torch.rand(M // rsd, device=dev) < 0.9).float() / 0.9) if rsd else None
    kw = dict(M=M, N=N, K=K, ldy=N, y_off=0, ldx=K, x_off=0, row_scale=rs, rs_div=rsd or 1, rs_binary=rsd > 0, dbias=db, ws=ws)
    res = []
    chunks = (M // rsd) * -(-rsd // 64) if rsd else -(-M // 64)
    for bn in (64, 128, 192, 256):
        if bn > max(64, (K + 63) // 64 * 64):
            continue
        tiles = -(-N // 128) * -(-K // bn)
        for ctas in (37, 74, 111, 148, 222, 296, 444, 592):
            parts = max(1, min(chunks, round(ctas / tiles)))
            if any(r[1] == bn and r[2] == parts for r in res):
                continue
            os.environ["CENET_B200_WGRAD_PLAN"] = f"{bn},{parts}"
            jobs, used = tops.gemm_wgrad_partial(dys[0], xs[0], dw, **kw)
            S = jobs[0][3] if jobs else 1
            t_r = 0.0
            if jobs:
                tab, nj, nb = tops.wgrad_reduce_table(jobs)
                tab = tab.to(dev)
                t_r = graph_time(lambda i: tops.wgrad_reduce_batch(tab, nj, nb), 4)
            t_p = graph_time(lambda i: tops.gemm_wgrad_partial(dys[i % R], xs[i % R], dw, **kw), 2 * R)
            res.append((t_p + t_r, bn, parts, S, t_p, t_r))
    os.environ.pop("CENET_B200_WGRAD_PLAN")
    out = (C.c_int * 9)()
    L.load().cenet_wgrad_plan_query(M, N, K, int(rsd > 0), rsd or 1, 1, 1 << 20, out)
    res.sort()
    mine = [r for r in res if r[1] == out[0] and r[3] == out[1]]
    print(f"{tag:9s} M={M} N={N} K={K} rs={rsd}: model bn={out[0]} S={out[1]}" + (f" ({mine[0][0]:.1f} us)" if mine else "") +
          " | best: " + "  ".join(f"bn={r[1]} S={r[3]} {r[4]:.1f}+{r[5]:.1f}" for r in res[:4]) +
          " | worst: " + f"bn={res[-1][1]} S={res[-1][3]} {res[-1][0]:.1f}", flush=True)
