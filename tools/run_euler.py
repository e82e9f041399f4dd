"""A few Euler-update launches on a 768 MiB working set (the bench's roofline_euler case) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
ctx = Context(0)
n = 32 * 1024 * 1024
x = torch.view_as_complex(torch.randn(n, 2, device="cuda")); v = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
for _ in range(4):
    ctx.euler_step(x, v, 0.2425)
torch.cuda.synchronize()
print("ok")
