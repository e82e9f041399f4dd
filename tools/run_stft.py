"""A ragged batch through the device STFT / iSTFT (for ncu): python tools/run_stft.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200.lib import Context
ctx = Context(0)
g = torch.Generator().manual_seed(0)
lens = [64000, 48000, 80000, 31000] * 8
wav = torch.zeros(len(lens), max(lens))
for i, L in enumerate(lens):
    wav[i, :L] = 0.1 * torch.randn(L, generator=g)
wav = wav.cuda()
for _ in range(3):
    Y, peak = ctx.stft_spec(wav, lens)
    out = ctx.spec_istft(Y, lens, peak=peak)
torch.cuda.synchronize()
print(Y.shape, float(out.abs().mean()))
