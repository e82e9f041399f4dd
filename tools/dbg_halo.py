"""Wait-cycle breakdown of the halo conv kernel (run with FLOWSE_CONV_DBG=1): python tools/dbg_halo.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
def split(a):
    hi = a.half(); return torch.stack([hi, (a - hi.float()).half()]).contiguous()
def run(H, W, Cin, Cout, impl):
    A = split(torch.randn(1, H, W, Cin, device="cuda", generator=g))
    w = torch.randn(Cout, Cin, 3, 3) / np.sqrt(Cin * 9)
    Wp, wexp = ctx.pack_conv_weights(w, None, ((Cout + 127) // 128) * 128)
    bias = torch.zeros(1, Cout, device="cuda")
    out = torch.empty(1, H, W, Cout, device="cuda")
    print(f"--- {H}x{W} Cin={Cin} Cout={Cout} impl={impl}", file=sys.stderr, flush=True)
    for _ in range(3):
        ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, out=out, impl=impl)
    torch.cuda.synchronize()
for impl in (2, 4):
    run(256, 512, 128, 128, impl)
    run(256, 512, 256, 128, impl)
    run(128, 256, 256, 256, impl)
