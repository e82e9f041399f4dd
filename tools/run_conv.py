"""Run the tcgen05 conv GEMM at a bench shape for ncu / timing: python tools/run_conv.py [Cin] [Cout] [H] [W] [Cin2] [iters]."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context

Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 128
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
W = int(sys.argv[4]) if len(sys.argv) > 4 else 512
Cin2 = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(1, H, W, Cin, device="cuda", generator=g)
hi = a.half(); A = torch.stack([hi, (a - hi.float()).half()]).contiguous()
X = None; wsc = None
if Cin2:
    x = torch.randn(1, H, W, Cin2, device="cuda", generator=g)
    xh = x.half(); X = torch.stack([xh, (x - xh.float()).half()]).contiguous()
    wsc = torch.randn(Cout, Cin2, 1, 1) / np.sqrt(Cin2)
w = torch.randn(Cout, Cin, 3, 3) / np.sqrt(Cin * 9)
npad = ((Cout + 127) // 128) * 128
Wp, wexp = ctx.pack_conv_weights(w, wsc, npad)
bias = torch.zeros(1, Cout, device="cuda")
res = torch.randn(1, H, W, Cout, device="cuda", generator=g) if not Cin2 else None
for _ in range(3):
    ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, X=X, residual=res, div_sqrt2=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
outs = []
e0.record()
for _ in range(iters):
    ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, X=X, residual=res, div_sqrt2=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
K = 9 * Cin + Cin2
fl = 2.0 * H * W * Cout * K
print(f"conv Cin={Cin} Cout={Cout} {H}x{W} Cin2={Cin2}: {ms*1e3:.1f} us/launch (incl. torch.zeros out alloc), {fl/ms/1e9:.1f} TFLOP/s algorithmic, {3*fl/ms/1e9:.1f} issued")
