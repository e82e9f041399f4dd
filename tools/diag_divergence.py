"""How fast do two valid fp32 evaluation orders diverge through long solver chains?  Compares the sampler output of
utterance 0 run alone (B=1) and inside a batch (B=4: other tile / split-K choices), and across conv accumulation
strategies, for Heun with N = 3, 5, 12, 25."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0))
def rc(shape, seed, s=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(s * torch.randn(*shape, 2, generator=g)).cuda()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Y, z = rc((4, 1, 256, T), 31, 0.3), rc((4, 1, 256, T), 32, np.sqrt(0.5))
def cmp(a, b):
    a, b = torch.view_as_real(a).cpu(), torch.view_as_real(b).cpu()
    d = (a - b).abs()
    return f"outside tol {((d > 1e-4 + 1e-3 * b.abs()).float().mean().item()):.3%}  max {d.max().item():.2e}  rms {d.pow(2).mean().sqrt().item():.2e}"
for N in (3, 5, 12, 25):
    ts = torch.linspace(1.0, 0.03, N)
    res = {}
    for impl in (0, 3, 2):
        ctx.set_option("conv_impl", impl)
        xb = ctx.sample(Y, z, ts, solver=1, sigma=0.487)
        x1 = ctx.sample(Y[:1].contiguous(), z[:1].contiguous(), ts, solver=1, sigma=0.487)
        res[impl] = (xb[:1].clone(), x1.clone())
        print(f"N={N:2d} Heun ({2*N-1} NFE) impl={impl}: batch-vs-single {cmp(xb[:1], x1)}", flush=True)
    print(f"N={N:2d} impl0 vs impl3 (B=1): {cmp(res[0][1], res[3][1])}")
    print(f"N={N:2d} impl0 vs impl2 (B=1): {cmp(res[0][1], res[2][1])}", flush=True)
