"""First-call cost of a new (B, T) plan vs steady state (plan build = arena allocation + eager run + graph capture)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context

ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0))
ts = torch.linspace(1.0, 0.03, 5)
for (B, T) in [(1, 512), (4, 256), (2, 512), (8, 128), (1, 512), (4, 256), (3, 192), (1, 1280)]:
    Y = torch.view_as_complex(0.3 * torch.randn(B, 1, 256, T, 2, device="cuda")); z = torch.randn_like(Y)
    torch.cuda.synchronize(); t0 = time.time()
    ctx.sample(Y, z, ts); torch.cuda.synchronize(); t1 = time.time()
    ctx.sample(Y, z, ts); torch.cuda.synchronize(); t2 = time.time()
    ctx.sample(Y, z, ts); torch.cuda.synchronize(); t3 = time.time()
    print(f"B={B} T={T}: first {1e3*(t1-t0):7.1f} ms, second {1e3*(t2-t1):7.1f} ms, steady {1e3*(t3-t2):7.1f} ms, workspace {ctx.workspace_bytes(B,T)/2**30:.2f} GiB")
