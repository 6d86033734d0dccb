"""Times the tcgen05 flash attention against the mma.sync kernel on the network's shapes:  python tools/one_attn_tc.py"""
import os
import subprocess
import sys
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run():
    import torch
    from cenet_b200 import ops
    dev = "cuda:0"
    res = {}
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    shapes = [("nonlocal dec1 C=64 N=3136", 64, 64, 1, 3136, 3136), ("nonlocal dec2 C=128 N=784", 64, 128, 1, 784, 784),
              ("sr enc1 C=64 N=3136 Nk=49", 64, 64, 1, 3136, 49), ("sr enc2 h=2 N=784", 64, 64, 2, 784, 49),
              ("sr enc3 h=5 N=196", 64, 64, 5, 196, 49), ("sr enc4 h=8 N=49", 64, 64, 8, 49, 49),
              ("nonlocal 512^2 C=64 N=16384", 16, 64, 1, 16384, 16384)]
    for name, B, D, heads, Nq, Nk in shapes:
        C = heads * D
        g = torch.Generator().manual_seed(0)
        if heads == 1 and Nq == Nk:
            tpg = torch.randn(B, Nq, 3 * C, generator=g).to(dev, torch.bfloat16)
            out = torch.empty(B, Nq, C, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.nonlocal_flash(tpg, out, B, Nq, C, C ** -0.5)
        else:
            q = torch.randn(B, Nq, C, generator=g).to(dev, torch.bfloat16)
            kv = torch.randn(B, Nk, 2 * C, generator=g).to(dev, torch.bfloat16)
            out = torch.empty(B, Nq, C, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.sr_attention(q, kv, out, B, Nq, Nk, C, heads, 0.125)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = 4.0 * Nq * Nk * D * heads * B
        res[name] = dict(ms=ms, tflops=fl / ms / 1e9, exps_per_clk_sm=Nq * Nk * heads * B / (ms * 1e-3) / (148 * 1.965e9))
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run()
    else:
        out = {}
        for tc in ("1", "0"):
            env = dict(os.environ, CENET_B200_ATTN_TC=tc)
            r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=env)
            try:
                out["tcgen05" if tc == "1" else "mma.sync"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                out["tcgen05" if tc == "1" else "mma.sync"] = dict(error=r.stderr[-2000:])
        for k in out.get("tcgen05", {}):
            a, b = out["tcgen05"].get(k), out.get("mma.sync", {}).get(k)
            print(f"{k:34s} tcgen05 {a}   mma.sync {b}")
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/attn_tc_vs_mma.json", "w"), indent=1)
