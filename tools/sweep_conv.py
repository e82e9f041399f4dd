"""K-sweep of the conv GEMM: time vs K at fixed tile count -> per-k-block cost and fixed per-tile overhead."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
def run(H, W, Cin, Cout, res=False, iters=20):
    a = torch.randn(1, H, W, Cin, device="cuda", generator=g)
    hi = a.half(); A = torch.stack([hi, (a - hi.float()).half()]).contiguous()
    w = torch.randn(Cout, Cin, 3, 3) / np.sqrt(Cin * 9)
    Wp, wexp = ctx.pack_conv_weights(w, None, ((Cout + 127) // 128) * 128)
    bias = torch.zeros(1, Cout, device="cuda")
    r = torch.randn(1, H, W, Cout, device="cuda", generator=g) if res else None
    out = torch.empty(1, H, W, Cout, device="cuda")
    for _ in range(3): ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, residual=r, div_sqrt2=True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, residual=r, div_sqrt2=True, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    tiles = (H * W // 128) * ((Cout + 127) // 128)
    kb = 9 * Cin // 64
    print(f"H={H} W={W} Cin={Cin} Cout={Cout} res={int(res)} tiles={tiles} kblocks={kb}: {us:8.1f} us  ({us/ -(-tiles//148):.2f} us/wave)")
for Cin in (64, 128, 256, 512):
    run(32, 512, Cin, 128)          # 128 tiles: one wave
for Cin in (64, 128, 256, 512):
    run(37*4, 128*4, Cin, 128)      # 592 tiles: 4 full waves
run(32, 512, 128, 128, res=True)
run(256, 512, 128, 128, res=False)
run(256, 512, 128, 128, res=True)
run(256, 512, 256, 128, res=False)
