"""Time the conv kernels op-level (CUDA events, preallocated output) for the dominant NCSN++ shapes:
python tools/bench_conv.py [impls, e.g. 0,2,3]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
impls = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,2,3").split(",")]
ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
def split(a):
    hi = a.half(); return torch.stack([hi, (a - hi.float()).half()]).contiguous()
def run(H, W, Cin, Cout, Cin2=0, res=False, iters=10):
    A = split(torch.randn(1, H, W, Cin, device="cuda", generator=g))
    X = split(torch.randn(1, H, W, Cin2, device="cuda", generator=g)) if Cin2 else None
    w = torch.randn(Cout, Cin, 3, 3) / np.sqrt(Cin * 9)
    wsc = torch.randn(Cout, Cin2, 1, 1) / np.sqrt(Cin2) if Cin2 else None
    npad = 16 if Cout < 16 else ((Cout + 127) // 128) * 128
    Wp, wexp = ctx.pack_conv_weights(w, wsc, npad)
    bias = torch.zeros(1, Cout, device="cuda")
    r = torch.randn(1, H, W, Cout, device="cuda", generator=g) if res else None
    out = torch.empty(1, H, W, Cout, device="cuda")
    K = 9 * Cin + Cin2
    fl = 2.0 * H * W * Cout * K
    line = f"{H:3d}x{W:3d} Cin={Cin:3d} Cout={Cout:3d} sc={Cin2:3d} res={int(res)} K={K:4d}:"
    ref = None
    for impl in impls:
        for _ in range(3): ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, X=X, residual=r, div_sqrt2=True, out=out, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, X=X, residual=r, div_sqrt2=True, out=out, impl=impl)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        if ref is None: ref = out.clone(); d = 0.0
        else: d = (out - ref).abs().max().item()
        line += f"  impl{impl} {us:7.1f} us {3*fl/us/1e6:6.0f} TF/s issued (d={d:.1e})"
    print(line, flush=True)
run(256, 512, 128, 128, res=True)
run(256, 512, 256, 128)
run(256, 512, 128, 128, Cin2=256)
run(128, 256, 128, 128, res=True)
run(128, 256, 256, 256, res=True)
run(128, 256, 256, 128, Cin2=384)
run(64, 128, 256, 256, res=True)
run(64, 128, 512, 256)
run(256, 512, 128, 4)
run(128, 256, 128, 4)
