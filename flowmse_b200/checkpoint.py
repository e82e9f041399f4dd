"""Checkpoint layout of the reference (Lightning 1.6.5 + torch_ema 0.3) and synthetic weights.

Reference layout (/root/reference/flowmse/model.py:81-90):
  ckpt['state_dict']        keys 'dnn.output_layer.*', 'dnn.all_modules.<i>.<...>'
  ckpt['ema']               torch_ema state: decay, num_updates, shadow_params (list in
                            model.parameters() order), collected_params
  ckpt['hyper_parameters']  backbone, ode, t_eps, T_rev, sigma_min, sigma_max, ...
Inference uses the EMA weights (evaluate.py:72 -> model.py:95-98).

Real FlowSE checkpoints are not available offline, and the reference's default
initialisation is degenerate (every Conv_1 / NIN_3 / pyramid conv is scaled by 1e-10,
layers.py:88-91 with init_scale=0, so the network output is a constant).  The
synthetic generator below therefore draws every tensor from a seeded, non-degenerate
distribution; parity tests and benchmarks use it on both sides.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import ncsnpp_spec as spec

DEFAULT_HPARAMS = dict(backbone="ncsnpp", ode="flowmatching", t_eps=0.03, T_rev=1.0,
                       sigma_min=0.0, sigma_max=0.487, lr=1e-4, ema_decay=0.999,
                       loss_type="mse", loss_abs_exponent=0.5, num_eval_files=10,
                       n_fft=510, hop_length=128, num_frames=256, window="hann",
                       spec_factor=0.15, spec_abs_exponent=0.5, transform_type="exponent")


def synthetic_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded non-degenerate NCSN++ weights in the reference's ``state_dict`` layout.

    Deterministic for a given torch version (CPU generator, one sub-seed per tensor).
    """
    sd: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(spec.state_dict_layout()):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 100003 + idx)
        leaf = name.rsplit(".", 1)[-1]
        if name == "all_modules.0.W":                       # GaussianFourierProjection (layerspp.py:37)
            t = torch.randn(shape, generator=g) * spec.FOURIER_SCALE
        elif len(shape) == 1 and leaf == "weight":          # GroupNorm scale (block-level or bare pyramid-head GN)
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("bias", "b"):                         # every bias / GroupNorm shift
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "W":                                   # NIN [in, out] (layers.py:546-549)
            t = torch.randn(shape, generator=g) / math.sqrt(shape[0])
        elif leaf == "weight":                              # conv / linear [out, in, ...]
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:
            raise AssertionError(name)
        sd[name] = t.to(torch.float32).contiguous()
    return sd


def parameter_names() -> List[str]:
    """Names in ``model.parameters()`` order == state-dict order (no buffers in NCSNpp)."""
    return [n for n, _ in spec.state_dict_layout()]


def make_lightning_checkpoint(sd: Dict[str, torch.Tensor], ema_sd: Optional[Dict[str, torch.Tensor]] = None,
                              hparams: Optional[dict] = None, include_frozen_in_ema: bool = True) -> dict:
    """Assemble a dict with the reference checkpoint layout (no Lightning needed).

    ``ema_sd`` are the EMA ("shadow") weights; defaults to ``sd``.  torch_ema 0.3 stores
    shadow params as a list in ``parameters()`` order; whether the frozen Fourier ``W``
    (requires_grad=False, layerspp.py:37) is included is version dependent, so the reader
    accepts both (647 or 646 entries).
    """
    ema_sd = sd if ema_sd is None else ema_sd
    names = parameter_names()
    if not include_frozen_in_ema:
        names = [n for n in names if n != "all_modules.0.W"]
    return {
        "state_dict": {"dnn." + k: v.clone() for k, v in sd.items()},
        "ema": {"decay": 0.999, "num_updates": 1, "shadow_params": [ema_sd[n].clone() for n in names],
                "collected_params": None},
        "hyper_parameters": dict(DEFAULT_HPARAMS if hparams is None else hparams),
        "pytorch-lightning_version": "1.6.5.post0",
        "epoch": 0, "global_step": 0,
    }


def backbone_state_from_checkpoint(ckpt: dict, use_ema: bool = True) -> Dict[str, torch.Tensor]:
    """Extract the NCSN++ weights the reference would run inference with.

    Mirrors VFModel.on_load_checkpoint + eval() (model.py:81-106): the EMA shadow
    parameters replace the live ones when present.
    """
    sd = {k[len("dnn."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("dnn.")}
    layout = spec.state_dict_layout()
    missing = [n for n, _ in layout if n not in sd]
    if missing:
        raise KeyError(f"checkpoint is missing {len(missing)} backbone tensors, e.g. {missing[:3]}")
    ema = ckpt.get("ema") if use_ema else None
    if ema is not None:
        shadow = ema["shadow_params"]
        names = parameter_names()
        if len(shadow) == len(names) - 1:
            names = [n for n in names if n != "all_modules.0.W"]
        elif len(shadow) != len(names):
            raise ValueError(f"EMA has {len(shadow)} shadow params, expected {len(names)} or {len(names) - 1}")
        for n, p in zip(names, shadow):
            if tuple(p.shape) != tuple(sd[n].shape):
                raise ValueError(f"EMA shadow param for {n} has shape {tuple(p.shape)}, expected {tuple(sd[n].shape)}")
            sd[n] = p
    out = {}
    for n, shape in layout:
        t = sd[n].detach().to(torch.float32).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{n}: shape {tuple(t.shape)} != {shape}")
        out[n] = t
    return out


def flatten_state_dict(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """One flat fp32 blob in state-dict order (what the C-ABI and the NCCL broadcast carry)."""
    return torch.cat([sd[n].reshape(-1).to(torch.float32) for n, _ in spec.state_dict_layout()])


def unflatten_state_dict(blob: torch.Tensor) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for n, shape in spec.state_dict_layout():
        k = 1
        for s in shape:
            k *= s
        out[n] = blob[off:off + k].reshape(shape)
        off += k
    assert off == blob.numel()
    return out


# ---------------------------------------------------------------------------------------------------------------
# Packed deployment checkpoints: the weights in libflowse's own layout (conv weights K-major fp16
# hi/lo pairs, exactly what the tcgen05 kernels read) plus the hyper-parameters the inference path needs.
# ---------------------------------------------------------------------------------------------------------------
PACKED_FORMAT = "flowse-packed-v1"


def save_packed(path: str, ctx, hparams: Optional[dict] = None) -> None:
    """Write the weights held by a libflowse ``Context`` (EMA weights after ``model.eval()``) as a packed checkpoint."""
    hp = dict(DEFAULT_HPARAMS)
    hp.update(hparams or {})
    hp.pop("data_module_cls", None)
    torch.save({"format": PACKED_FORMAT, "hyper_parameters": hp, "blob": ctx.export_packed()}, path)


def load_packed(path: str) -> dict:
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ck, dict) or ck.get("format") != PACKED_FORMAT:
        raise ValueError(f"{path} is not a {PACKED_FORMAT} checkpoint")
    return ck
