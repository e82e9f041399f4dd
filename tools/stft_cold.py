"""Where the first flowse_stft_spec call of a process spends its time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
t0 = time.time(); ctx = Context(0); torch.cuda.synchronize(); print(f"context {time.time()-t0:.3f} s")
for L, B in ((131072, 1), (131072, 1), (98304, 3), (65536, 8), (65536, 8)):
    wav = torch.randn(B, L, device="cuda"); torch.cuda.synchronize()
    t0 = time.time(); Y, peak = ctx.stft_spec(wav, [L] * B); torch.cuda.synchronize(); t1 = time.time()
    x = ctx.spec_istft(Y, [L] * B, peak=peak); torch.cuda.synchronize(); t2 = time.time()
    print(f"B={B} L={L}: stft_spec {1e3*(t1-t0):8.2f} ms, spec_istft {1e3*(t2-t1):8.2f} ms")
