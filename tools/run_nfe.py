"""Run a few NCSN++ evaluations at the bench shape (B=1, T=512) for ncu: `python tools/run_nfe.py [n] [graph]`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
graph = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
T = int(sys.argv[4]) if len(sys.argv) > 4 else 512
ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0)); ctx.set_option("graph", graph)
g = torch.Generator().manual_seed(0)
xy = torch.view_as_complex(0.3 * torch.randn(B, 2, 256, T, 2, generator=g)).cuda()
t = torch.full((B,), 0.515, device="cuda")
for _ in range(n):
    v = ctx.ncsnpp_forward(xy, t)
torch.cuda.synchronize()
print("launches", ctx.kernel_launches(), float(v.abs().mean()))
