"""CPU oracle for the FlowSE reverse-ODE sampling hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the CPU baseline.  The product package
(``flowmse_b200``) never imports it and fails loudly without its CUDA library.

What this is: a restatement, in plain fp32 torch-CPU tensor ops, of the
reference's algorithm for the path named by BASELINE.json's north_star.  Every
function cites the reference file:line it follows (paths relative to
/root/reference).  It consumes a plain ``state_dict`` in the reference's own key
layout (``all_modules.<i>.<...>`` / ``output_layer.*`` as produced by
``flowmse/backbones/ncsnpp.py:99-245``), so it also pins the checkpoint layout.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/gen_golden.py`` imports /root/reference (with stubs for the
absent pytorch_lightning / torch_ema / matplotlib / pesq / pystoi), runs the
reference's NCSNpp.forward, upfirdn2d and get_white_box_solver on seeded inputs
and commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against those fixtures on every CPU run.
Heun / midpoint are NOT in the reference source (only in a stale .pyc); their
formulas here are the survey's disassembly and are "parity
unpinned" - pinned only through the reference's own VFModel.forward driven by
the restated update rule.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------
# Architecture walk (default config of flowmse/backbones/ncsnpp.py:45-67)
# ----------------------------------------------------------------------------
NF = 128
CH_MULT = (1, 1, 2, 2, 2, 2, 2)
NUM_RES_BLOCKS = 2
ATTN_RESOLUTIONS = (16,)
IMAGE_SIZE = 256
NUM_CHANNELS = 4  # x.re, x.im, y.re, y.im (ncsnpp.py:95)


def build_module_list() -> List[dict]:
    """Re-derive the flat ``all_modules`` list (ncsnpp.py:99-245), as dicts.

    kinds: fourier, linear, conv3x3, rb (ResnetBlockBigGANpp), attn, combine, gn.
    """
    mods: List[dict] = []
    nf = NF
    num_res = len(CH_MULT)
    all_res = [IMAGE_SIZE // (2 ** i) for i in range(num_res)]
    mods.append(dict(kind="fourier", size=nf))                      # ncsnpp.py:102-105
    mods.append(dict(kind="linear", cin=2 * nf, cout=4 * nf))       # :114
    mods.append(dict(kind="linear", cin=4 * nf, cout=4 * nf))       # :117
    mods.append(dict(kind="conv3x3", cin=NUM_CHANNELS, cout=nf))    # :163
    hs_c = [nf]
    in_ch = nf
    for i_level in range(num_res):                                   # :167-195
        for _ in range(NUM_RES_BLOCKS):
            out_ch = nf * CH_MULT[i_level]
            mods.append(dict(kind="rb", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if all_res[i_level] in ATTN_RESOLUTIONS:
                mods.append(dict(kind="attn", c=in_ch))
            hs_c.append(in_ch)
        if i_level != num_res - 1:
            mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=True))
            mods.append(dict(kind="combine", cin=NUM_CHANNELS, cout=in_ch))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]                                                 # :197-200
    mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False))
    mods.append(dict(kind="attn", c=in_ch))
    mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False))
    for i_level in reversed(range(num_res)):                         # :204-236
        for _ in range(NUM_RES_BLOCKS + 1):
            out_ch = nf * CH_MULT[i_level]
            mods.append(dict(kind="rb", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if all_res[i_level] in ATTN_RESOLUTIONS:
            mods.append(dict(kind="attn", c=in_ch))
        mods.append(dict(kind="gn", c=in_ch))
        mods.append(dict(kind="conv3x3", cin=in_ch, cout=NUM_CHANNELS))
        if i_level != 0:
            mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    return mods


# ----------------------------------------------------------------------------
# Primitive restatements
# ----------------------------------------------------------------------------
def silu(x: torch.Tensor) -> torch.Tensor:
    """nn.SiLU (layers.py:38-39)."""
    return x * torch.sigmoid(x)


def _fir_up_axis(x: torch.Tensor, dim: int) -> torch.Tensor:
    """One axis of upsample_2d(x,[1,3,3,1],factor=2) (up_or_down_sampling.py:195-224):
    zero-stuff x2, taps [1,3,3,1]/8*2, pad (2,1) => out[2m]=.25x[m-1]+.75x[m],
    out[2m+1]=.75x[m]+.25x[m+1], zero boundary."""
    n = x.shape[dim]
    xm1 = torch.roll(x, 1, dim)
    xp1 = torch.roll(x, -1, dim)
    idx0 = [slice(None)] * x.ndim
    idx0[dim] = slice(0, 1)
    xm1[tuple(idx0)] = 0
    idxl = [slice(None)] * x.ndim
    idxl[dim] = slice(n - 1, n)
    xp1[tuple(idxl)] = 0
    even = 0.25 * xm1 + 0.75 * x
    odd = 0.75 * x + 0.25 * xp1
    out = torch.stack([even, odd], dim=dim + 1)
    shape = list(x.shape)
    shape[dim] = 2 * n
    return out.reshape(shape)


def fir_upsample2(x: torch.Tensor) -> torch.Tensor:
    """upsample_2d(x, (1,3,3,1), factor=2) for NCHW x (up_or_down_sampling.py:195-224,
    op/upfirdn2d.py:159-200).  Separable polyphase form."""
    return _fir_up_axis(_fir_up_axis(x, 2), 3)


def _fir_down_axis(x: torch.Tensor, dim: int) -> torch.Tensor:
    """One axis of downsample_2d(x,[1,3,3,1],factor=2) (up_or_down_sampling.py:227-257):
    pad (1,1), taps [1,3,3,1]/8, stride 2 => out[m]=(x[2m-1]+3x[2m]+3x[2m+1]+x[2m+2])/8."""
    n = x.shape[dim]
    pad = [0, 0] * x.ndim
    # F.pad pads from the last dim backwards
    pad[2 * (x.ndim - 1 - dim)] = 1
    pad[2 * (x.ndim - 1 - dim) + 1] = 1
    xp = F.pad(x, pad)

    def sl(start):
        idx = [slice(None)] * x.ndim
        idx[dim] = slice(start, start + n - 1, 2)
        return xp[tuple(idx)]

    return (sl(0) + 3.0 * sl(1) + 3.0 * sl(2) + sl(3)) / 8.0


def fir_downsample2(x: torch.Tensor) -> torch.Tensor:
    """downsample_2d(x, (1,3,3,1), factor=2) for NCHW x (up_or_down_sampling.py:227-257)."""
    return _fir_down_axis(_fir_down_axis(x, 2), 3)


def group_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.GroupNorm(min(C//4,32), C, eps=1e-6) (layerspp.py:219,231,67; ncsnpp.py:210)."""
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), w, b, eps=1e-6)


def nin(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """NIN.forward (layers.py:546-555): channel mix y = x.W + b with W [in, out]."""
    y = torch.einsum("bchw,cd->bdhw", x, W)
    return y + b[None, :, None, None]


def fourier_embedding(t: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """GaussianFourierProjection(log t) (layerspp.py:39-41, ncsnpp.py:259)."""
    x = torch.log(t)
    x_proj = x[:, None] * W[None, :] * 2 * np.pi
    return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


def resblock(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, up: bool, down: bool) -> torch.Tensor:
    """ResnetBlockBigGANpp.forward (layerspp.py:242-274)."""
    cin = x.shape[1]
    cout = sd[p + "Conv_0.weight"].shape[0]
    h = silu(group_norm(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"]))
    if up:
        h = fir_upsample2(h)
        x = fir_upsample2(x)
    elif down:
        h = fir_downsample2(h)
        x = fir_downsample2(x)
    h = F.conv2d(h, sd[p + "Conv_0.weight"], sd[p + "Conv_0.bias"], padding=1)
    h = h + F.linear(silu(temb), sd[p + "Dense_0.weight"], sd[p + "Dense_0.bias"])[:, :, None, None]
    h = silu(group_norm(h, sd[p + "GroupNorm_1.weight"], sd[p + "GroupNorm_1.bias"]))
    h = F.conv2d(h, sd[p + "Conv_1.weight"], sd[p + "Conv_1.bias"], padding=1)
    if cin != cout or up or down:
        x = F.conv2d(x, sd[p + "Conv_2.weight"], sd[p + "Conv_2.bias"])
    return (x + h) / np.sqrt(2.0)


def attnblock(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AttnBlockpp.forward (layerspp.py:75-91)."""
    B, C, H, W = x.shape
    h = group_norm(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"])
    q = nin(h, sd[p + "NIN_0.W"], sd[p + "NIN_0.b"])
    k = nin(h, sd[p + "NIN_1.W"], sd[p + "NIN_1.b"])
    v = nin(h, sd[p + "NIN_2.W"], sd[p + "NIN_2.b"])
    w = torch.einsum("bchw,bcij->bhwij", q, k) * (int(C) ** (-0.5))
    w = F.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
    h = torch.einsum("bhwij,bcij->bchw", w, v)
    h = nin(h, sd[p + "NIN_3.W"], sd[p + "NIN_3.b"])
    return (x + h) / np.sqrt(2.0)


# ----------------------------------------------------------------------------
# Backbone
# ----------------------------------------------------------------------------
def ncsnpp_forward(sd: SD, x: torch.Tensor, time_cond: torch.Tensor,
                   taps: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """NCSNpp.forward (ncsnpp.py:247-404).

    x: complex64 [B,2,F,T] (channel 0 = state, 1 = noisy condition), time_cond fp32 [B].
    Returns complex64 [B,1,F,T].  If ``taps`` is a dict, per-module activations are
    stored under ``m<idx>`` for layer-wise parity tests.
    """
    mods = build_module_list()
    m = 0

    def pre(i):
        return f"all_modules.{i}."

    def tap(i, v):
        if taps is not None:
            taps[f"m{i}"] = v

    x4 = torch.cat((x[:, [0]].real, x[:, [0]].imag, x[:, [1]].real, x[:, [1]].imag), dim=1)  # :253-254
    temb = fourier_embedding(time_cond, sd[pre(0) + "W"]); m += 1                              # :259
    temb = F.linear(temb, sd[pre(1) + "weight"], sd[pre(1) + "bias"]); m += 1                  # :272
    temb = F.linear(silu(temb), sd[pre(2) + "weight"], sd[pre(2) + "bias"]); m += 1           # :274
    tap(2, temb)
    input_pyramid = x4
    hs = [F.conv2d(x4, sd[pre(3) + "weight"], sd[pre(3) + "bias"], padding=1)]; tap(3, hs[0]); m += 1  # :285
    num_res = len(CH_MULT)
    for i_level in range(num_res):                                                             # :289-322
        for _ in range(NUM_RES_BLOCKS):
            h = resblock(sd, pre(m), hs[-1], temb, False, False); tap(m, h); m += 1
            if h.shape[-2] in ATTN_RESOLUTIONS:
                h = attnblock(sd, pre(m), h); tap(m, h); m += 1
            hs.append(h)
        if i_level != num_res - 1:
            h = resblock(sd, pre(m), hs[-1], temb, False, True); tap(m, h); m += 1
            input_pyramid = fir_downsample2(input_pyramid)                                     # :310
            h = F.conv2d(input_pyramid, sd[pre(m) + "Conv_0.weight"], sd[pre(m) + "Conv_0.bias"]) + h  # Combine, layerspp.py:52-57
            tap(m, h); m += 1
            hs.append(h)
    h = hs[-1]
    h = resblock(sd, pre(m), h, temb, False, False); tap(m, h); m += 1                         # :324-330
    h = attnblock(sd, pre(m), h); tap(m, h); m += 1
    h = resblock(sd, pre(m), h, temb, False, False); tap(m, h); m += 1
    pyramid = None
    for i_level in reversed(range(num_res)):                                                   # :335-385
        for _ in range(NUM_RES_BLOCKS + 1):
            h = resblock(sd, pre(m), torch.cat([h, hs.pop()], dim=1), temb, False, False); tap(m, h); m += 1
        if h.shape[-2] in ATTN_RESOLUTIONS:
            h = attnblock(sd, pre(m), h); tap(m, h); m += 1
        ph = silu(group_norm(h, sd[pre(m) + "weight"], sd[pre(m) + "bias"])); m += 1           # :347-366
        ph = F.conv2d(ph, sd[pre(m) + "weight"], sd[pre(m) + "bias"], padding=1); m += 1
        pyramid = ph if pyramid is None else fir_upsample2(pyramid) + ph
        tap(m - 1, pyramid)
        if i_level != 0:
            h = resblock(sd, pre(m), h, temb, True, False); tap(m, h); m += 1                  # :379-385
    assert not hs and m == len(mods)
    h = pyramid / time_cond[:, None, None, None]                                               # :398
    tap(1000, pyramid)
    h = F.conv2d(h, sd["output_layer.weight"], sd["output_layer.bias"])                        # :401
    h = h.permute(0, 2, 3, 1).contiguous()
    return torch.view_as_complex(h)[:, None, :, :]                                             # :402-403


def vf_forward(sd: SD, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """VFModel.forward (model.py:164-170): -dnn(cat([x, y], 1), t)."""
    return -ncsnpp_forward(sd, torch.cat([x, y], dim=1), t)


# ----------------------------------------------------------------------------
# Sampler
# ----------------------------------------------------------------------------
def schedule(N: int, T_rev: float = 1.0, t_eps: float = 0.03) -> Tuple[torch.Tensor, torch.Tensor]:
    """timesteps / stepsizes of get_white_box_solver (sampling/__init__.py:45-53), fp32."""
    ts = torch.linspace(T_rev, t_eps, N)
    steps = torch.empty(N, dtype=torch.float32)
    for i in range(N):
        steps[i] = ts[i] - ts[i + 1] if i != N - 1 else ts[-1]
    return ts, steps


def prior_sample(y: torch.Tensor, z: torch.Tensor, sigma_min: float = 0.0, sigma_max: float = 0.487) -> torch.Tensor:
    """FLOWMATCHING.prior_sampling with the noise z passed in (odes.py:86-100)."""
    t1 = torch.ones((y.shape[0],))
    std = (1 - t1) * sigma_min + t1 * sigma_max
    return y + z * std[:, None, None, None]


def sample(sd: SD, Y: torch.Tensor, z: torch.Tensor, N: int, solver: str = "euler",
           T_rev: float = 1.0, t_eps: float = 0.03, sigma_min: float = 0.0, sigma_max: float = 0.487,
           vf=None) -> torch.Tensor:
    """get_white_box_solver(...)() (sampling/__init__.py:27-62) with z given.

    solver: 'euler' (odesolvers.py:37-47); 'heun' / 'midpoint' per SURVEY section 8 A4
    (unpinned; last step falls back to Euler because VF(., t+dt=0) is NaN: ncsnpp.py:259,398).
    """
    if vf is None:
        vf = lambda x, t, y: vf_forward(sd, x, t, y)
    with torch.no_grad():
        xt = prior_sample(Y, z, sigma_min, sigma_max)
        ts, _ = schedule(N, T_rev, t_eps)
        for i in range(N):
            t = ts[i]
            stepsize = t - ts[i + 1] if i != N - 1 else ts[-1]
            vec_t = torch.ones(Y.shape[0]) * t
            dt = -stepsize
            last = i == N - 1
            if solver == "euler" or last:
                xt = xt + vf(xt, vec_t, Y) * dt
            elif solver == "heun":
                v0 = vf(xt, vec_t, Y)
                x_next = xt + dt * v0
                xt = xt + dt / 2 * (v0 + vf(x_next, vec_t + dt, Y))
            elif solver == "midpoint":
                x_mid = xt + dt / 2 * vf(xt, vec_t, Y)
                xt = xt + dt * vf(x_mid, vec_t + dt / 2, Y)
            else:
                raise ValueError(f"ODEsolver with name '{solver}' unknown.")
        return xt


def pad_spec(Y: torch.Tensor) -> torch.Tensor:
    """pad_spec (util/other.py:83-90): zero-pad T to a multiple of 64."""
    T = Y.size(3)
    num_pad = 64 - T % 64 if T % 64 != 0 else 0
    return F.pad(Y, (0, num_pad, 0, 0))


# ----------------------------------------------------------------------------
# STFT / iSTFT either side of the sampler
# ----------------------------------------------------------------------------
def stft_spec(wav: torch.Tensor, normalize: bool = True, spec_factor: float = 0.15, e: float = 0.5):
    """evaluate.py:109-115 for one utterance [1, L]: peak-normalise, SpecsDataModule.stft (data_module.py:163-170),
    spec_fwd (data_module.py:149-162), pad_spec (util/other.py:83-90).  Returns (Y [1,1,256,Tpad], peak)."""
    peak = wav.abs().max() if normalize else torch.tensor(1.0)
    win = torch.hann_window(510, periodic=True)
    S = torch.stft(wav / peak, n_fft=510, hop_length=128, window=win, center=True, return_complex=True)
    S = S.abs() ** e * torch.exp(1j * S.angle()) * spec_factor
    return pad_spec(S.unsqueeze(0)), peak


def spec_istft(X: torch.Tensor, length: int, peak=1.0, spec_factor: float = 0.15, e: float = 0.5) -> torch.Tensor:
    """VFModel.to_audio (model.py:190-203): spec_back (data_module.py:164-175) then torch.istft(..., length), times the
    peak (evaluate.py:134-135).  X: [1,1,256,T] or [256,T]."""
    S = X.squeeze() / spec_factor
    S = S.abs() ** (1 / e) * torch.exp(1j * S.angle())
    win = torch.hann_window(510, periodic=True)
    return torch.istft(S, n_fft=510, hop_length=128, window=win, center=True, length=length) * peak
