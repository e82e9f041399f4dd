"""Probability path of FlowSE: mirror of /root/reference/flowmse/odes.py (ODERegistry, FLOWMATCHING).

Only what the sampling path touches is device-accelerated (prior_sampling -> libflowse prior kernel on CUDA
tensors); the training-side helpers are kept as plain tensor expressions for API completeness.
"""
import abc
import warnings

import torch

from .util.registry import Registry

ODERegistry = Registry("ODE")


class ODE(abc.ABC):
    @abc.abstractmethod
    def marginal_prob(self, x, t, *args):
        ...

    @abc.abstractmethod
    def prior_sampling(self, shape, *args):
        ...

    @staticmethod
    @abc.abstractmethod
    def add_argparse_args(parent_parser):
        ...

    @abc.abstractmethod
    def copy(self):
        ...


@ODERegistry.register("flowmatching")
class FLOWMATCHING(ODE):
    """mu_t = (1-t) x0 + t y, sigma_t = (1-t) sigma_min + t sigma_max (odes.py:59-107)."""

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--sigma_min", type=float, default=0.00)
        parser.add_argument("--sigma_max", type=float, default=0.487)
        return parser

    def __init__(self, sigma_min=0.00, sigma_max=0.487, **ignored_kwargs):
        super().__init__()
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max

    def copy(self):
        return FLOWMATCHING(self.sigma_min, self.sigma_max)

    def ode(self, x, t, *args):
        pass

    def _mean(self, x0, t, y):
        return (1 - t)[:, None, None, None] * x0 + t[:, None, None, None] * y

    def _std(self, t):
        return (1 - t) * self.sigma_min + t * self.sigma_max

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_std(self) -> float:
        """sigma(t=1) as an fp32-rounded Python float: (1-1)*sigma_min + 1*sigma_max evaluated like odes.py:86-88,96."""
        t1 = torch.ones((1,))
        return float(((1 - t1) * self.sigma_min + t1 * self.sigma_max)[0])

    def prior_sampling(self, shape, y):
        """x_T = y + sigma(1) z with z ~ randn_like(y) from torch's generator of y's device (odes.py:93-100)."""
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        z = torch.randn_like(y)
        from .runtime import get_context
        x_T = get_context(y.device).prior_sample(y.contiguous(), z, self.prior_std())
        return x_T, z

    def der_mean(self, x0, t, y):
        return y - x0

    def der_std(self, t):
        return self.sigma_max - self.sigma_min
