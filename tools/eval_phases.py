"""Where the evaluate drop-in's wall time goes on a cold process: per-phase timers around enhance_batch's steps for the
utterances one rank of an 8-GPU run would get (103 of the 824 synthetic utterances)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmse_b200 import evaluate as ev, sharding
from flowmse_b200.checkpoint import synthetic_state_dict
from flowmse_b200.model import VFModel

t00 = time.time()
model = VFModel(backbone="ncsnpp", ode="flowmatching"); model.dnn.load_state_dict(synthetic_state_dict(0), strict=True); model.eval()
dev = torch.device("cuda", 0)
pairs = ev.synthetic_test_set(824)
n_samples = [len(p[2]) for p in pairs]
mine = sharding.lpt_assign([ev.padded_frames(n) for n in n_samples], 8)[0]
batches = ev.make_batches(mine, n_samples, 4096)
ctx = model.flowse_context(dev)
torch.cuda.synchronize(); print(f"setup (weights, context): {time.time()-t00:.2f} s, {len(mine)} files in {len(batches)} batches")
tot = dict(h2d=0.0, stft=0.0, sampler=0.0, istft=0.0)
frames = 0
for rep in range(2):
    for k in tot: tot[k] = 0.0
    rows = []
    for batch in batches:
        t0 = time.time()
        wavs = [torch.from_numpy(pairs[i][2]).to(dev) for i in batch]
        lens = [int(w.numel()) for w in wavs]
        Tpad = ev.padded_frames(max(lens))
        wav = torch.zeros((len(batch), max(lens)), device=dev)
        for r, w in enumerate(wavs): wav[r, :lens[r]] = w
        torch.cuda.synchronize(); t1 = time.time()
        Y, peak = ctx.stft_spec(wav, lens, Tpad=Tpad)
        torch.cuda.synchronize(); t2 = time.time()
        X = model.enhance_spec(Y, N=5)
        torch.cuda.synchronize(); t3 = time.time()
        xh = ctx.spec_istft(X.contiguous(), lens, peak=peak)
        torch.cuda.synchronize(); t4 = time.time()
        tot["h2d"] += t1 - t0; tot["stft"] += t2 - t1; tot["sampler"] += t3 - t2; tot["istft"] += t4 - t3
        rows.append((len(batch), Tpad, round(1e3 * (t3 - t2), 1), "stft", round(1e3 * (t2 - t1), 1), "istft", round(1e3 * (t4 - t3), 1)))
    frames = sum(ev.padded_frames(n_samples[i]) for i in mine)
    print(("cold" if rep == 0 else "warm"), {k: round(v, 3) for k, v in tot.items()}, "frames", frames,
          "frames/s (sampler only)", round(frames / tot["sampler"]))
    if rep == 0: print("cold per batch (B, T, sampler ms):", rows)
    else: print("warm per batch (B, T, sampler ms):", rows)
