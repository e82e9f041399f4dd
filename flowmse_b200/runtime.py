"""Per-device libflowse contexts shared by the drop-in classes."""
from __future__ import annotations

from typing import Dict

import torch

from .lib import Context, FlowseError

_contexts: Dict[int, Context] = {}


def get_context(device) -> Context:
    """The process-wide context of a CUDA device (created on first use).  Raises on CPU tensors: no fallback."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise FlowseError(f"flowmse_b200 runs on CUDA (B200) tensors only, got device '{dev}'; there is no CPU fallback")
    idx = torch.cuda.current_device() if dev.index is None else dev.index
    ctx = _contexts.get(idx)
    if ctx is None:
        with torch.cuda.device(idx):
            ctx = Context(idx)
        _contexts[idx] = ctx
    return ctx


def new_context(device) -> Context:
    """A private context (own weights + workspace); used when several backbones live on one device."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise FlowseError(f"flowmse_b200 runs on CUDA (B200) tensors only, got device '{dev}'; there is no CPU fallback")
    idx = torch.cuda.current_device() if dev.index is None else dev.index
    with torch.cuda.device(idx):
        return Context(idx)
