"""A few sampler calls (for ncu): python tools/run_sample.py [B] [T] [N]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
N = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0))
g = torch.Generator().manual_seed(0)
Y = torch.view_as_complex(0.3 * torch.randn(B, 1, 256, T, 2, generator=g)).cuda()
z = torch.view_as_complex(torch.randn(B, 1, 256, T, 2, generator=g) * 0.5 ** 0.5).cuda()
ctx.set_option("graph", 0)
for _ in range(2):
    x = ctx.sample(Y, z, torch.linspace(1.0, 0.03, N))
torch.cuda.synchronize()
print(float(x.abs().mean()))
