import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cenet_b200 import ops
B, N, E, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
qkv = torch.randn(B, N, 3 * E, device="cuda").to(torch.bfloat16)
out = torch.empty(B, N, E, device="cuda", dtype=torch.bfloat16)
ws = torch.empty(B * 2 * heads, device="cuda") if os.environ.get("BOUNDED", "1") == "1" else None
for _ in range(3):
    ops.diffattn_flash(qkv, out, B, N, E, heads, 0.5, 1e-5, 0.5, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.diffattn_flash(qkv, out, B, N, E, heads, 0.5, 1e-5, 0.5, ws)
e1.record(); torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 5)
