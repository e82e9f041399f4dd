"""Batch-vs-single consistency at B=16: python tools/diag_batch.py T"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0))
def rc(shape, seed, s=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(s * torch.randn(*shape, 2, generator=g)).cuda()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Y, z = rc((B, 1, 256, T), 31, 0.3), rc((B, 1, 256, T), 32, np.sqrt(0.5))
def cmp(a, b):
    a, b = torch.view_as_real(a).cpu(), torch.view_as_real(b).cpu()
    d = (a - b).abs()
    return f"outside tol {((d > 1e-4 + 1e-3 * b.abs()).float().mean().item()):.3%}  max {d.max().item():.2e}  rms {d.pow(2).mean().sqrt().item():.2e}"
for solver, N in ((0, 1), (0, 5), (1, 5), (1, 25)):
    ts = torch.linspace(1.0, 0.03, N)
    for impl in (0, 3, 2):
        ctx.set_option("conv_impl", impl)
        xb = ctx.sample(Y, z, ts, solver=solver, sigma=0.487)
        for i in (0, B - 5):
            x1 = ctx.sample(Y[i:i+1].contiguous(), z[i:i+1].contiguous(), ts, solver=solver, sigma=0.487)
            print(f"T={T} B={B} solver={solver} N={N:2d} impl={impl} elem {i:2d}: {cmp(xb[i:i+1], x1)}", flush=True)
