"""eval_phases' first batch with finer timers (one-off stall hunt)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.model import VFModel
def T(msg, t0):
    torch.cuda.synchronize(); t1 = time.time(); print(f"{msg:44s} {1e3*(t1-t0):9.2f} ms", flush=True); return time.time()
t0 = time.time()
model = VFModel(backbone="ncsnpp", ode="flowmatching"); model.dnn.load_state_dict(synthetic_state_dict(0), strict=True); model.eval()
dev = torch.device("cuda", 0)
ctx = model.flowse_context(dev); t0 = T("model + context", t0)
for (B, L) in ((1, 131000), (1, 98000), (3, 90000)):
    wav = torch.randn(B, L, device=dev); t0 = T(f"randn wav B={B}", t0)
    Y, peak = ctx.stft_spec(wav, [L] * B); t0 = T("stft_spec", t0)
    X = model.enhance_spec(Y, N=5); t0 = T("enhance_spec", t0)
    Xc = X.contiguous(); t0 = T("contiguous", t0)
    xh = ctx.spec_istft(Xc, [L] * B, peak=peak); t0 = T("spec_istft", t0)
    y2 = xh * 2; t0 = T("torch mul on result", t0)
