"""Per-op device times of one NCSN++ evaluation (B=1, T=512) for the conv layers, fused operand prep off / on:
python tools/xf_diag.py  (env FLOWSE_HALO_LAYOUT selects the warp-role layout)."""
import os, sys
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
res = {}
for fuse in (0, 1):
    ctx.set_option("fuse_prep", fuse)
    for _ in range(3):
        ctx.ncsnpp_forward(xy, t)
    torch.cuda.synchronize()
    ctx.profile_forward()
    ops = ctx.profile_forward()
    res[fuse] = ops
    tot = sum(o["ms"] for o in ops)
    fam = {}
    for o in ops:
        fam[o["kind"]] = fam.get(o["kind"], 0.0) + o["ms"]
    print(f"fuse={fuse}: {len(ops)} ops, sum {tot:.3f} ms, by family {{" + ", ".join(f"{k}: {v:.3f}" for k, v in fam.items()) + "}")
c0 = [o for o in res[0] if o["kind"] in ("conv_halo", "conv_gemm")]
c1 = [o for o in res[1] if o["kind"] in ("conv_halo", "conv_gemm")]
print("conv layers (H, W, K, Cout): us unfused -> fused")
agg = {}
for a, b in zip(c0, c1):
    if a["kind"] != "conv_halo":
        continue
    key = (a["H"], a["W"], a["K"], a["Cout"])
    e = agg.setdefault(key, [0, 0.0, 0.0])
    e[0] += 1; e[1] += a["ms"]; e[2] += b["ms"]
for key, (n, m0, m1) in sorted(agg.items(), reverse=True):
    fl = 2.0 * B * key[0] * key[1] * key[2] * key[3]
    print(f"  {key} x{n}: {1e3*m0/n:7.1f} -> {1e3*m1/n:7.1f} us  ({fl/(m0/n*1e-3)/1e12:.0f} -> {fl/(m1/n*1e-3)/1e12:.0f} TFLOP/s alg)")
