"""The probability path the sampler integrates: drop-in for /root/reference/flowmse/odes.py (ODERegistry, FLOWMATCHING).

FlowSE's conditional flow-matching path is Gaussian with mean mu_t = (1-t) x0 + t y and scale
sigma_t = (1-t) sigma_min + t sigma_max (odes.py:83-91); sampling starts from x_1 = y + sigma_1 z (odes.py:93-100).
On this path only ``prior_sampling`` runs on the device (libflowse ``flowse_prior_sample``); the remaining methods are the
closed forms the reference's training code calls, kept so that code written against the reference class keeps working.
"""
from __future__ import annotations

import warnings

import torch

from .util.registry import Registry

ODERegistry = Registry("ODE")

_DEFAULTS = {"sigma_min": 0.00, "sigma_max": 0.487}          # odes.py:67-68,71


def _per_sample(v: torch.Tensor) -> torch.Tensor:
    """[B] -> [B,1,1,1] so a per-utterance scalar broadcasts over a [B,1,F,T] spectrogram."""
    return v.reshape(-1, 1, 1, 1)


class ODE:
    """Interface the sampler and the model expect from an entry of ODERegistry (odes.py:18-52): ``marginal_prob``,
    ``prior_sampling``, ``copy`` and the static ``add_argparse_args``.  Subclasses override all four."""

    def marginal_prob(self, x0, t, y):
        raise NotImplementedError

    def prior_sampling(self, shape, y):
        raise NotImplementedError

    def copy(self):
        raise NotImplementedError

    @staticmethod
    def add_argparse_args(parser):
        return parser


@ODERegistry.register("flowmatching")
class FLOWMATCHING(ODE):
    def __init__(self, sigma_min=_DEFAULTS["sigma_min"], sigma_max=_DEFAULTS["sigma_max"], **ignored_kwargs):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    @staticmethod
    def add_argparse_args(parser):
        for flag, default in _DEFAULTS.items():
            parser.add_argument(f"--{flag}", type=float, default=default)
        return parser

    def copy(self):
        return type(self)(sigma_min=self.sigma_min, sigma_max=self.sigma_max)

    # ---- closed forms of the path (odes.py:81-91,102-107) ------------------------------------------------------
    def ode(self, x, t, *args):
        return None                                   # a stub in the reference as well (odes.py:81-82)

    def _std(self, t):
        return (1 - t) * self.sigma_min + t * self.sigma_max

    def _mean(self, x0, t, y):
        return _per_sample(1 - t) * x0 + _per_sample(t) * y

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def der_mean(self, x0, t, y):
        return y - x0

    def der_std(self, t):
        return self.sigma_max - self.sigma_min

    # ---- the part on the sampling path --------------------------------------------------------------------------
    def prior_std(self) -> float:
        """sigma_1 exactly as the reference's fp32 tensor arithmetic produces it: ``_std(torch.ones(B))`` (odes.py:86-88,96)."""
        return float(self._std(torch.ones(1))[0])

    def prior_sampling(self, shape, y):
        """(x_1, z) with x_1 = y + sigma_1 z.  z comes from ``torch.randn_like(y)``, i.e. from the global generator of y's
        device, which is what keeps "same seed, same sample" against the reference (odes.py:97)."""
        if tuple(shape) != tuple(y.shape):
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        from .runtime import get_context
        z = torch.randn_like(y)
        return get_context(y.device).prior_sample(y.contiguous(), z, self.prior_std()), z
