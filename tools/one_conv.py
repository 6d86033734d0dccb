"""Times the OutHead convolutions (B=64) with the halo-resident conv mode on / off:  python tools/one_conv.py"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from cenet_b200 import ops
    dev = "cuda:0"
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    out = {}
    for name, B, Cin, Cout, k, H in [("rb.conv2 5x5 32->32 @224", 64, 32, 32, 5, 224), ("out.0 3x3 64->64 @112", 64, 64, 64, 3, 112),
                                     ("up 3x3 64->32 @112", 64, 64, 32, 3, 112), ("5x5 32->32 @512 B16", 16, 32, 32, 5, 512)]:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, H, H, Cin, generator=g).to(dev, torch.bfloat16)
        w = (torch.randn(Cout, k * k * Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev, torch.bfloat16)
        bias = torch.randn(Cout, generator=g).to(dev)
        res = torch.randn(B * H * H, Cout, generator=g).to(dev, torch.bfloat16)
        o = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.conv_nhwc(x, w, o, k, 1, k // 2, bias=bias, act=ops.ACT_LEAKY, slope=0.01, act_after_res=True, res1=res, ldr1=Cout)
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
        fl = 2.0 * B * H * H * Cout * k * k * Cin
        by = (B * H * H * (Cin + 2 * Cout)) * 2
        out[name] = dict(ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), gbs=round(by / ms / 1e6, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    res = {}
    for halo in ("1", "0"):
        r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=dict(os.environ, CENET_B200_CONV_HALO=halo))
        try:
            res["halo" if halo == "1" else "shifted boxes"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res[halo] = dict(error=(r.stdout + r.stderr)[-800:])
    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/conv_halo.json", "w"), indent=1)
