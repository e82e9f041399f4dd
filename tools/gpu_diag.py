"""Layer-by-layer diagnosis on the GPU box: CUDA taps vs full oracle activations (CPU, computed here)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
from oracle import ncsnpp_oracle as orc

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = 2
sd = synthetic_state_dict(0)
g = torch.Generator().manual_seed(21)
x = torch.view_as_complex(0.3 * torch.randn(B, 2, 256, T, 2, generator=g))
t = torch.tensor([0.757, 0.03])
taps = {}
with torch.no_grad():
    v_ref = orc.ncsnpp_forward(sd, x, t, taps=taps)
ctx = Context(0); ctx.load_state_dict(sd); ctx.set_option("graph", 0); ctx.set_option("conv_impl", impl)
v = ctx.ncsnpp_forward(x.cuda(), t.cuda()).cpu()
rows = []
for k in sorted(taps, key=lambda s: int(s[1:])):
    m = int(k[1:])
    if m in (2, 1000): continue
    try:
        tap = ctx.debug_tap(m, B).cpu()
    except Exception as e:
        continue
    ref = taps[k]
    d = (tap - ref).abs()
    rows.append(dict(m=m, shape=list(ref.shape), max_err=d.max().item(), rel=d.max().item() / ref.abs().max().item(),
                     rms_rel=(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()))
    print(rows[-1])
d = torch.view_as_real(v - v_ref).abs()
print("OUT max err", d.max().item(), "ref max", v_ref.abs().max().item(), "per batch", [d[b].max().item() for b in range(B)])
bad = d > 1e-4 + 1e-3 * torch.view_as_real(v_ref).abs()
print("outside tol:", bad.float().mean().item())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/diag_T{T}_impl{impl}.json", "w"))
