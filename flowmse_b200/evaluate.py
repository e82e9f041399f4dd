"""Batched, length-bucketed, multi-GPU drop-in for the reference's evaluate.py.

Same command line and the same output files as /root/reference/evaluate.py:27-43,161-194 (``files/*.wav``,
``_results.csv``, ``_avg_results.txt``, ``_settings.txt``); what changes is how the work is done:

  reference (evaluate.py:97-136)                      here
  ------------------------------------------------    ---------------------------------------------------------------
  one file at a time, `.item()` sync for the peak     files are grouped by padded frame count (pad_spec buckets) and run
  torch.stft on one utterance                         as batches: ONE device STFT, ONE sampler call, ONE device iSTFT
  N Python-level solver steps                         per batch (VFModel.enhance_batch -> flowse_stft_spec /
  one GPU                                             flowse_sample / flowse_spec_istft)
                                                      one process per GPU (torchrun): files are LPT-assigned to ranks by
                                                      frame count, every rank writes its own wav files, the metric rows
                                                      are gathered on rank 0

    python -m flowmse_b200.evaluate --test_dir DATA --folder_destination OUT --ckpt CKPT --N 5
    torchrun --nproc-per-node 8 -m flowmse_b200.evaluate ...            (utterance-sharded)

Metrics: SI-SDR / SI-SIR / SI-SAR restate /root/reference/utils.py:10-36; PESQ and ESTOI come from the third-party
``pesq`` / ``pystoi`` packages exactly as in the reference when they are installed and are NaN otherwise (they are CPU
code outside the hot path).  Without a dataset, ``--synthetic_utts K`` builds the VoiceBank-DEMAND-shaped synthetic test
set of BASELINE.json configs[4] (lengths from a fixed-seed log-normal) so the whole driver can be exercised
offline; ``--synthetic_weights SEED`` replaces ``--ckpt`` by the seeded non-degenerate weights used everywhere else.
"""
from __future__ import annotations

import glob
import json
import math
import os
import re
import time
from argparse import ArgumentParser
from os.path import join
from typing import Dict, List, Sequence

import numpy as np
import torch

SR = 16000


# ---------------------------------------------------------------------------------------------------------------
# metrics (utils.py:10-36) and the small helpers evaluate.py imports from utils.py
# ---------------------------------------------------------------------------------------------------------------
def energy_ratios(s_hat, s, n):
    """Scale-invariant SDR / SIR / SAR in dB (Le Roux et al. 2019; same decomposition as the reference's utils.py:10-36):
    the estimate is projected onto the clean signal ``s`` and onto the noise ``n``; what is left is the artefact term."""
    s_hat, s, n = (np.asarray(v, dtype=np.float64) for v in (s_hat, s, n))
    target = s * (s_hat @ s / (s @ s))
    noise = n * (s_hat @ n / (n @ n))
    artefact = s_hat - target - noise
    energy = lambda v: float(v @ v)
    db = lambda num, den: 10.0 * np.log10(num / den)
    e_t = energy(target)
    return db(e_t, energy(noise + artefact)), db(e_t, energy(noise)), db(e_t, energy(artefact))


def print_mean_std(data, decimals=2):
    data = np.array(data, dtype=np.float64)
    data = data[~np.isnan(data)]
    if data.size == 0:
        return "nan ± nan"
    return f"{np.mean(data):.{decimals}f} ± {np.std(data):.{decimals}f}"


def _optional_metric_fns():
    try:
        from pesq import pesq as _pesq
    except Exception:
        _pesq = None
    try:
        from pystoi import stoi as _stoi
    except Exception:
        _stoi = None
    return _pesq, _stoi


def file_metrics(x: np.ndarray, y: np.ndarray, x_hat: np.ndarray) -> Dict[str, float]:
    """The per-file metric row of evaluate.py:147-158 (x clean, y noisy, x_hat enhanced)."""
    _pesq, _stoi = _optional_metric_fns()
    n = y - x
    try:
        p = _pesq(SR, x, x_hat, "wb") if _pesq else float("nan")
    except Exception:
        p = float("nan")
    e = _stoi(x, x_hat, SR, extended=True) if _stoi else float("nan")
    sdr, sir, sar = energy_ratios(x_hat, x, n)
    return dict(pesq=float(p), estoi=float(e), si_sdr=float(sdr), si_sir=float(sir), si_sar=float(sar))


# ---------------------------------------------------------------------------------------------------------------
# wav I/O (scipy; the reference uses torchaudio.load / soundfile.write, neither is on the arithmetic path)
# ---------------------------------------------------------------------------------------------------------------
def read_wav(path: str) -> np.ndarray:
    from scipy.io import wavfile
    sr, a = wavfile.read(path)
    if sr != SR:
        raise ValueError(f"{path}: sample rate {sr}, expected {SR}")
    if a.ndim > 1:
        a = a[:, 0]
    if a.dtype == np.int16:
        return (a.astype(np.float32) / 32768.0)
    if a.dtype == np.int32:
        return (a.astype(np.float32) / 2147483648.0)
    return a.astype(np.float32)


def write_wav(path: str, x: np.ndarray):
    """16-bit PCM, what soundfile.write(path, x, 16000) produces for a .wav (evaluate.py:144)."""
    from scipy.io import wavfile
    wavfile.write(path, SR, np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16))


# ---------------------------------------------------------------------------------------------------------------
# work partitioning
# ---------------------------------------------------------------------------------------------------------------
def frames_of(n_samples: int) -> int:
    return 1 + int(n_samples) // 128


def padded_frames(n_samples: int) -> int:
    return ((frames_of(n_samples) + 63) // 64) * 64


def make_batches(indices: Sequence[int], n_samples: Sequence[int], max_batch_frames: int) -> List[List[int]]:
    """Group utterances of identical padded frame count into batches of at most `max_batch_frames` total frames
    (always at least one utterance per batch), longest bucket first so the big allocations happen once."""
    by_t: Dict[int, List[int]] = {}
    for i in indices:
        by_t.setdefault(padded_frames(n_samples[i]), []).append(i)
    out: List[List[int]] = []
    for t in sorted(by_t, reverse=True):
        per = max(1, max_batch_frames // t)
        idx = by_t[t]
        out += [idx[k:k + per] for k in range(0, len(idx), per)]
    return out


def synthetic_test_set(k: int, seed: int = 0):
    """K synthetic (clean, noisy) pairs with VoiceBank-DEMAND-like durations: log-normal, median 2.4 s, clipped to
    [1, 10] s (SURVEY.md 8d config 5 - the real histogram is not in the reference, this is the stated assumption)."""
    rng = np.random.RandomState(seed)
    secs = np.clip(np.exp(rng.normal(math.log(2.4), 0.45, size=k)), 1.0, 10.0)
    pairs = []
    for j, s in enumerate(secs):
        n = int(round(s * SR))
        t = np.arange(n) / SR
        f0 = 110.0 + 30.0 * (j % 7)
        clean = (0.35 * np.sin(2 * np.pi * f0 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 3.1 * t)) +
                 0.15 * np.sin(2 * np.pi * 3.03 * f0 * t)).astype(np.float32)
        noisy = (clean + 0.08 * rng.standard_normal(n)).astype(np.float32)
        pairs.append((f"synth_{j:04d}.wav", clean, noisy))
    return pairs


# ---------------------------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------------------------
def build_parser() -> ArgumentParser:
    p = ArgumentParser()
    # the reference's arguments (evaluate.py:27-43), same names, defaults and choices
    p.add_argument("--test_dir", type=str, default=None, help="Directory containing the test data")
    p.add_argument("--odesolver_type", type=str, choices=("white",), default="white")
    p.add_argument("--odesolver", type=str, choices=("euler", "heun", "midpoint"), default="euler", help="Numerical integrator")
    p.add_argument("--reverse_starting_point", type=float, default=1.0, help="Starting point for the ODE.")
    p.add_argument("--last_eval_point", type=float, default=0.03)
    p.add_argument("--folder_destination", type=str, required=True, help="Destination path of inference results.")
    p.add_argument("--ckpt", type=str, default=None, help="Path to model checkpoint.")
    p.add_argument("--N", type=int, default=5, help="Number of time steps")
    p.add_argument("--N_mid", type=int, default=0, help="It is not related to FlowSE")
    # additions
    p.add_argument("--max_batch_frames", type=int, default=4096, help="frames per sampler batch (B * padded T)")
    p.add_argument("--synthetic_utts", type=int, default=0, help="use K synthetic utterances instead of --test_dir")
    p.add_argument("--synthetic_weights", type=int, default=None, help="seeded synthetic weights instead of --ckpt")
    p.add_argument("--seed", type=int, default=None, help="torch.manual_seed(seed + rank) before sampling")
    return p


def load_model(args):
    from .model import VFModel
    if args.ckpt:
        model = VFModel.load_from_checkpoint(args.ckpt, base_dir="", batch_size=8, num_workers=4, kwargs=dict(gpu=False))
    elif args.synthetic_weights is not None:
        from .checkpoint import synthetic_state_dict
        model = VFModel(backbone="ncsnpp", ode="flowmatching")
        model.dnn.load_state_dict(synthetic_state_dict(args.synthetic_weights), strict=True)
    else:
        raise SystemExit("need --ckpt or --synthetic_weights")
    model.eval(no_ema=False)
    return model


def main(argv=None) -> Dict[str, float]:
    import torch.distributed as dist
    from . import sharding

    args = build_parser().parse_args(argv)
    if args.N_mid != 0:
        raise ValueError("N_mid should be 0.")
    if not torch.cuda.is_available():
        raise SystemExit("flowmse_b200.evaluate needs a CUDA device; there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    # ---- weights: rank 0 reads the checkpoint, ONE broadcast of the flat fp32 blob ---------------------------------
    from .checkpoint import flatten_state_dict, unflatten_state_dict
    model = load_model(args) if (rank == 0 or world == 1) else None
    if world > 1:
        from . import ncsnpp_spec
        from .model import VFModel
        if rank == 0:
            blob = flatten_state_dict(model.dnn.state_dict()).to(dev)
            # every hyper-parameter the enhancement depends on: all ranks must apply the same STFT / transform
            hp = [dict(t_eps=model.t_eps, T_rev=model.T_rev, sigma_min=model.ode.sigma_min, sigma_max=model.ode.sigma_max,
                       **model.data_module.hparams())]
        else:
            blob = torch.empty(ncsnpp_spec.num_params(), dtype=torch.float32, device=dev)
            hp = [None]
        dist.broadcast_object_list(hp, src=0)
        sharding.broadcast_weights(blob)
        if rank != 0:
            model = VFModel(backbone="ncsnpp", ode="flowmatching", **hp[0])
            model.dnn.load_state_dict(unflatten_state_dict(blob.cpu()), strict=True)
            model.eval()
        del blob
    model.T_rev = args.reverse_starting_point
    model.t_eps = args.last_eval_point

    # ---- the file list (identical on every rank) ----------------------------------------------------------------
    if args.synthetic_utts > 0:
        pairs = synthetic_test_set(args.synthetic_utts)
        names = [p[0] for p in pairs]
        load_pair = lambda i: (pairs[i][1], pairs[i][2])
        n_samples = [len(p[2]) for p in pairs]
    else:
        if not args.test_dir:
            raise SystemExit("need --test_dir or --synthetic_utts")
        clean_dir, noisy_dir = join(args.test_dir, "test", "clean"), join(args.test_dir, "test", "noisy")
        noisy_files = sorted(glob.glob("{}/*.wav".format(noisy_dir)))
        names = [f.split("/")[-1] for f in noisy_files]
        load_pair = lambda i: (read_wav(join(clean_dir, names[i])), read_wav(noisy_files[i]))
        from scipy.io import wavfile   # header-only length scan
        n_samples = [int(wavfile.read(f, mmap=True)[1].shape[0]) for f in noisy_files]

    target_dir = f"{args.folder_destination}/"
    os.makedirs(target_dir + "files/", exist_ok=True)
    mine = sharding.lpt_assign([padded_frames(n) for n in n_samples], world)[rank]
    batches = make_batches(mine, n_samples, args.max_batch_frames)
    # One-off start-up, kept out of the per-file timing like the reference's `model.cuda()` (evaluate.py:72-73): device
    # context, weight upload + packing, kernel loading, STFT bases - exercised by a 1 s dummy utterance.  Reported
    # separately (`startup_seconds`); done before the seed is set so that results for a given --seed do not depend on it.
    t_start = time.time()
    if dev.type == "cuda":
        model.enhance_batch([torch.zeros(16000, device=dev) + 1e-3], N=1, odesolver=args.odesolver)
        torch.cuda.synchronize()
    startup = time.time() - t_start
    if args.seed is not None:
        torch.manual_seed(args.seed + rank)

    rows, frames_done, t_gpu = [], 0, 0.0
    for batch in batches:
        data = [load_pair(i) for i in batch]
        wavs = [torch.from_numpy(d[1]).to(dev) for d in data]
        torch.cuda.synchronize()
        t0 = time.time()
        enhanced = model.enhance_batch(wavs, N=args.N, odesolver=args.odesolver)
        torch.cuda.synchronize()
        t_gpu += time.time() - t0
        for i, (x, y), xh in zip(batch, data, enhanced):
            x_hat = xh.cpu().numpy()
            write_wav(target_dir + "files/" + names[i], x_hat)
            m = min(len(x), len(x_hat))
            rows.append(dict(filename=names[i], **file_metrics(x[:m], y[:m], x_hat[:m])))
            frames_done += padded_frames(len(y))

    stats = dict(rank=rank, files=len(rows), frames=frames_done, seconds=t_gpu, startup_seconds=startup)
    if world > 1:
        all_rows, all_stats = [None] * world, [None] * world
        dist.all_gather_object(all_rows, rows)
        dist.all_gather_object(all_stats, stats)
        rows = [r for part in all_rows for r in part]
        stats_list = all_stats
    else:
        stats_list = [stats]
    summary = {}
    if rank == 0:
        import pandas as pd
        rows.sort(key=lambda r: r["filename"])
        df = pd.DataFrame(rows, columns=["filename", "pesq", "estoi", "si_sdr", "si_sir", "si_sar"])
        df.to_csv(join(target_dir, "_results.csv"), index=False)
        with open(join(target_dir, "_avg_results.txt"), "w") as f:
            f.write("PESQ: {} \n".format(print_mean_std(df["pesq"])))
            f.write("ESTOI: {} \n".format(print_mean_std(df["estoi"])))
            f.write("SI-SDR: {} \n".format(print_mean_std(df["si_sdr"])))
            f.write("SI-SIR: {} \n".format(print_mean_std(df["si_sir"])))
            f.write("SI-SAR: {} \n".format(print_mean_std(df["si_sar"])))
        match = re.search(r"epoch=(\d+)", args.ckpt or "")
        with open(join(target_dir, "_settings.txt"), "w") as f:
            f.write(f"epoch: {match.group(1) if match else 'n/a'}\n")
            f.write("checkpoint file: {}\n".format(args.ckpt))
            f.write("odesolver_type: {}\n".format(args.odesolver_type))
            f.write("odesolver: {}\n".format(args.odesolver))
            f.write("Reverse starting point: {}\n".format(args.reverse_starting_point))
            f.write("Last evaluated point: {}\n".format(args.last_eval_point))
            f.write("data: {}\n".format(args.test_dir))
            f.write("ode: {}\n".format("FLOWMATCHING"))
            f.write(f"sigma_min: {model.ode.sigma_min}\n")
            f.write(f"sigma_max: {model.ode.sigma_max}\n")
            f.write("N: {}\n".format(args.N))
        slowest = max(s["seconds"] for s in stats_list)
        total_frames = sum(s["frames"] for s in stats_list)
        summary = dict(files=len(rows), frames=total_frames, n_gpus=world, seconds_max_rank=slowest,
                       startup_seconds_max_rank=max(s.get("startup_seconds", 0.0) for s in stats_list),
                       frames_per_s=total_frames / slowest if slowest > 0 else float("nan"), per_rank=stats_list,
                       N=args.N, odesolver=args.odesolver)
        with open(join(target_dir, "_timing.json"), "w") as f:
            json.dump(summary, f, indent=1)
        print(json.dumps({k: v for k, v in summary.items() if k != "per_rank"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return summary


if __name__ == "__main__":
    main()
