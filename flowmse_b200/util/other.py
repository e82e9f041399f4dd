"""Shape contract of the hot path: pad_spec (/root/reference/flowmse/util/other.py:83-90)."""
import torch
import torch.nn.functional as F


def pad_spec(Y: torch.Tensor) -> torch.Tensor:
    """Zero-pad the frame axis (dim 3) up to a multiple of 64."""
    T = Y.size(3)
    num_pad = 64 - T % 64 if T % 64 != 0 else 0
    return F.pad(Y, (0, num_pad, 0, 0))
