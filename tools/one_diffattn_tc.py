"""Bring-up + timing of the tcgen05 differential attention:  python tools/one_diffattn_tc.py [check|time]
check: error vs a torch fp32 reference for every descriptor hypothesis (CENET_DA_TC_DESC 0..3) and the mma.sync kernel.
time : S56 (B=64, E=128, 8 heads), S28, and the 512^2 skin level for tcgen05 (poly 0/1/2) and mma.sync."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ref(qkv, B, N, E, heads, lam, mult):
    import torch
    hd = E // heads // 2
    q = qkv[..., :E].view(B, N, 2 * heads, hd).transpose(1, 2) * hd ** -0.5
    k = qkv[..., E:2 * E].view(B, N, 2 * heads, hd).transpose(1, 2)
    v = qkv[..., 2 * E:].view(B, N, heads, 2 * hd).transpose(1, 2)
    s = torch.softmax(q @ k.transpose(-1, -2), -1).view(B, heads, 2, N, N)
    o = (s[:, :, 0] - lam * s[:, :, 1]) @ v
    o = o * torch.rsqrt(o.pow(2).mean(-1, keepdim=True) + 1e-5) * mult
    return o.transpose(1, 2).reshape(B, N, E)


def child(mode):
    import torch
    from cenet_b200 import ops
    dev = "cuda:0"
    out = {}
    if mode == "check":
        for E, heads, N in [(128, 8, 200), (256, 8, 130), (256, 4, 257), (256, 2, 130), (128, 8, 1000)]:
            B = 2
            g = torch.Generator().manual_seed(E + heads)
            qkv = torch.randn(B, N, 3 * E, generator=g)
            qkv[..., :2 * E] *= 2.0
            qb = qkv.to(dev, torch.bfloat16)
            r = ref(qb.float(), B, N, E, heads, 0.55, 0.45)
            for use_kmax in (0, 1):
                o = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
                ws = torch.empty(B * 2 * heads, device=dev) if use_kmax else None
                ops.diffattn_flash(qb, o, B, N, E, heads, 0.55, 1e-5, 0.45, ws)
                torch.cuda.synchronize()
                out[f"E{E}_h{heads}_N{N}_kmax{use_kmax}"] = ((o.float() - r).norm() / r.norm()).item()
    else:
        flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
        for name, B, E, heads, N in [("S56 synapse hd8", 64, 128, 8, 3136), ("S28 synapse hd16", 64, 256, 8, 784),
                                     ("S56 acdc hd16", 24, 128, 4, 3136), ("S128 skin512 hd32", 16, 128, 2, 16384),
                                     ("S64 skin512 hd64", 16, 256, 2, 4096)]:
            g = torch.Generator().manual_seed(0)
            qb = torch.randn(B, N, 3 * E, generator=g).to(dev, torch.bfloat16)
            o = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
            ws = torch.empty(B * 2 * heads, device=dev)
            fn = lambda: ops.diffattn_flash(qb, o, B, N, E, heads, 0.55, 1e-5, 0.45, ws)
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            out[name] = dict(ms=round(ms, 4), tflops=round(4.0 * N * N * E * B / ms / 1e9, 1),
                             exps_per_clk_sm=round(2.0 * heads * N * N * B / (ms * 1e-3) / (148 * 1.965e9), 2))
    print(json.dumps(out))


def run(mode, env):
    r = subprocess.run([sys.executable, __file__, "child", mode], capture_output=True, text=True, env=dict(os.environ, **env), timeout=600)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return dict(error=(r.stdout + r.stderr)[-600:])


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
        sys.exit(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    res = {}
    if what == "check":
        res["mma.sync"] = run("check", dict(CENET_B200_DIFFATTN_TC="0"))
        for dm in ("0", "1", "2", "3"):
            res[f"tc desc{dm}"] = run("check", dict(CENET_DA_TC_DESC=dm))
    else:
        res["mma.sync"] = run("time", dict(CENET_B200_DIFFATTN_TC="0"))
        for pp in ("0", "1", "2", "3"):
            res[f"tc poly{pp}"] = run("time", dict(CENET_DA_TC_POLY=pp))
    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/diffattn_tc_{what}.json", "w"), indent=1)
