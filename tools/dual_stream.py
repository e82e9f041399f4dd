"""Throughput of two half-batches on two streams / two contexts vs one batch on one stream (prep of one half overlaps the
tensor-bound convs of the other): python tools/dual_stream.py [B_total] [T]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sd = synthetic_state_dict(0)
c0, c1, c2 = Context(0), Context(0), Context(0)
for c in (c0, c1, c2):
    c.load_state_dict(sd)
ts = torch.linspace(1.0, 0.03, 5)
Y = torch.view_as_complex(0.3 * torch.randn(B, 1, 256, T, 2, device="cuda")); z = torch.randn_like(Y)
Ya, Yb = Y[:B // 2].contiguous(), Y[B // 2:].contiguous(); za, zb = z[:B // 2].contiguous(), z[B // 2:].contiguous()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def one():
    return c0.sample(Y, z, ts)

def two():
    with torch.cuda.stream(s1):
        xa = c1.sample(Ya, za, ts)
    with torch.cuda.stream(s2):
        xb = c2.sample(Yb, zb, ts)
    return xa, xb

for fn, name in ((one, "one stream, B=%d" % B), (two, "two streams, 2 x B=%d" % (B // 2))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 5
    print(f"{name}: {dt*1e3:.1f} ms per pass, {B*T/dt:.0f} frames/s")
x1 = one(); xa, xb = two(); torch.cuda.synchronize()
print("max diff", (x1[:B // 2] - xa).abs().max().item(), (x1[B // 2:] - xb).abs().max().item())
