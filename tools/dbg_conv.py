import os, sys
os.environ["FLOWSE_CONV_DBG"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
ctx = Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
def run(H, W, Cin, Cout, res=False):
    a = torch.randn(1, H, W, Cin, device="cuda", generator=g)
    hi = a.half(); A = torch.stack([hi, (a - hi.float()).half()]).contiguous()
    w = torch.randn(Cout, Cin, 3, 3) / np.sqrt(Cin * 9)
    Wp, wexp = ctx.pack_conv_weights(w, None, ((Cout + 127) // 128) * 128)
    bias = torch.zeros(1, Cout, device="cuda")
    r = torch.randn(1, H, W, Cout, device="cuda", generator=g) if res else None
    out = torch.empty(1, H, W, Cout, device="cuda")
    for _ in range(3): ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, residual=r, div_sqrt2=True, out=out)
    torch.cuda.synchronize()
run(32, 512, 128, 128); run(32, 512, 512, 128); run(256, 512, 128, 128); run(256, 512, 128, 128, True)
# low-resolution layers (cluster split-K through distributed shared memory)
print("--- low-res (cluster split-K)", file=sys.stderr)
run(4, 8, 256, 256); run(8, 16, 256, 256); run(16, 32, 256, 256); run(16, 32, 512, 256); run(32, 64, 256, 256); run(64, 128, 256, 256)
