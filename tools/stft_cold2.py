"""Find the one-off ~0.5-1 s stall seen around the first STFT / iSTFT call after a sampler call (tools/eval_phases.py)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.lib import Context
def T(msg, t0):
    torch.cuda.synchronize(); t1 = time.time(); print(f"{msg:40s} {1e3*(t1-t0):9.2f} ms"); return time.time()
t0 = time.time(); ctx = Context(0); ctx.load_state_dict(synthetic_state_dict(0)); t0 = T("context + weights", t0)
ts = torch.linspace(1.0, 0.03, 5)
L = 131072
wav = torch.randn(1, L, device="cuda"); t0 = T("randn wav", t0)
Y, peak = ctx.stft_spec(wav, [L]); t0 = T("stft_spec (1, 1024)", t0)
z = torch.randn_like(Y); t0 = T("randn_like", t0)
X = ctx.sample(Y, z, ts); t0 = T("sample (1, 1024) first", t0)
Xc = X.contiguous(); t0 = T("contiguous", t0)
out = torch.empty((1, L), device="cuda"); t0 = T("torch.empty", t0)
x = ctx.spec_istft(Xc, [L], peak=peak); t0 = T("spec_istft (1, 1024)", t0)
x = ctx.spec_istft(Xc, [L], peak=peak); t0 = T("spec_istft again", t0)
wav = torch.randn(3, 90000, device="cuda"); t0 = T("randn wav", t0)
Y, peak = ctx.stft_spec(wav, [90000] * 3); t0 = T("stft_spec (3, 704)", t0)
X = ctx.sample(Y, torch.randn_like(Y), ts); t0 = T("sample (3, 704) first", t0)
x = ctx.spec_istft(X.contiguous(), [90000] * 3, peak=peak); t0 = T("spec_istft (3, 704)", t0)
