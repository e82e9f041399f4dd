"""Per-op device-time table of one NFE at the bench shape (CUDA events after every op)."""
import os, sys, json, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0))
g = torch.Generator().manual_seed(0)
xy = torch.view_as_complex(0.3 * torch.randn(B, 2, 256, T, 2, generator=g)).cuda()
t = torch.full((B,), 0.515, device="cuda")
for _ in range(2): ctx.ncsnpp_forward(xy, t)
ctx.profile_forward()
ops = ctx.profile_forward()
tot = sum(o["ms"] for o in ops)
agg = collections.OrderedDict()
for o in ops:
    key = (o["kind"], o["H"], o["W"], o["K"], o["Cout"]) if o["kind"] == "conv_gemm" else (o["kind"],)
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += o["ms"]; a[2] += o["flops"]
print(f"total {tot:.3f} ms over {len(ops)} ops")
for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f" {fl/ms/1e9:7.1f} TF/s alg, {3*fl/ms/1e9:7.1f} issued" if fl else ""
    print(f"{str(k):45s} n={n:3d} {ms*1e3:9.1f} us  avg {ms*1e3/n:7.1f}{extra}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(ops, open(f"gpurun_out/ops_B{B}_T{T}.json", "w"))
