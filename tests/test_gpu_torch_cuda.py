"""configs[1] against the torch-CUDA evaluation of the oracle (the stand-in for "reference torch-cuda": same ATen ops the
reference dispatches - cuDNN convolutions, native_group_norm, bmm, softmax - driven by the restated module walk).

Two things are checked / recorded on the B200:
  * parity: our N=5 Euler sampler output vs the fp32 torch-CUDA path (TF32 disabled) within rtol 1e-3 / atol 1e-4;
  * a timing record (CUDA events) of that torch path with TF32 off and on, written to gpurun_out/torch_cuda_baseline.json
    so that profiles/ can quote "x times the GPU-PyTorch sampler" from a measurement rather than a guess.  The TF32 run
    also reports how many bins leave the tolerance (why single-pass reduced precision is not an option).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ncsnpp_oracle as orc

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand_c(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(scale * torch.randn(*shape, 2, generator=g))


def _torch_cuda_sample(sd_cuda, Y, z, N, reps):
    """oracle.sample with every tensor on cuda:0; returns (x, ms per sampler call)."""
    with torch.device("cuda"):
        x = orc.sample(sd_cuda, Y, z, N)                 # warm-up: cuDNN heuristics, allocator
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            x = orc.sample(sd_cuda, Y, z, N)
        e1.record()
        torch.cuda.synchronize()
    return x, e0.elapsed_time(e1) / reps


def test_config2_vs_torch_cuda(synthetic_sd):
    from flowmse_b200.lib import Context
    T, N = 512, 5
    Y, z = _rand_c((1, 1, 256, T), 0, 0.3).cuda(), _rand_c((1, 1, 256, T), 1234, np.sqrt(0.5)).cuda()
    sd_cuda = {k: v.cuda() for k, v in synthetic_sd.items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    rec = {"workload": f"B=1, 2x256x{T}, N={N} Euler", "gpu": torch.cuda.get_device_name(0)}
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        x_fp32, ms_fp32 = _torch_cuda_sample(sd_cuda, Y, z, N, reps=3)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        x_tf32, ms_tf32 = _torch_cuda_sample(sd_cuda, Y, z, N, reps=3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    c = Context(0)
    try:
        c.load_state_dict(synthetic_sd)
        ts = torch.linspace(1.0, 0.03, N)
        for _ in range(3):
            x = c.sample(Y, z, ts, solver=0, sigma=0.487)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            x = c.sample(Y, z, ts, solver=0, sigma=0.487)
        e1.record()
        torch.cuda.synchronize()
        ms_ours = e0.elapsed_time(e1) / 5
    finally:
        c.close()

    def frac_bad(a, b):
        a, b = torch.view_as_real(a), torch.view_as_real(b)
        return ((a - b).abs() > ATOL + RTOL * b.abs()).float().mean().item(), (a - b).abs().max().item()

    f_ours, mx_ours = frac_bad(x, x_fp32)
    f_tf32, mx_tf32 = frac_bad(x_tf32, x_fp32)
    rec.update({
        "torch_cuda_fp32_ms": ms_fp32, "torch_cuda_fp32_frames_per_s": T / (ms_fp32 * 1e-3),
        "torch_cuda_tf32_ms": ms_tf32, "torch_cuda_tf32_frames_per_s": T / (ms_tf32 * 1e-3),
        "flowse_ms": ms_ours, "flowse_frames_per_s": T / (ms_ours * 1e-3),
        "speedup_vs_torch_cuda_fp32": ms_fp32 / ms_ours, "speedup_vs_torch_cuda_tf32": ms_tf32 / ms_ours,
        "flowse_vs_fp32": {"frac_outside_tol": f_ours, "max_abs": mx_ours},
        "tf32_vs_fp32": {"frac_outside_tol": f_tf32, "max_abs": mx_tf32},
    })
    print("torch-cuda baseline:", json.dumps(rec))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "torch_cuda_baseline.json"), "w") as f:
            json.dump(rec, f, indent=1)
    except OSError:
        pass
    assert f_ours == 0.0, (f_ours, mx_ours)
