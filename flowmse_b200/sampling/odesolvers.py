"""Step operators: mirror of /root/reference/flowmse/sampling/odesolvers.py (ODEsolverRegistry, ODEsolver,
EulerODEsolver) plus the Heun / midpoint rules of SURVEY.md section 8 A4.

``update_fn(x, t, y, stepsize)`` works with ANY ``VF_fn`` (the plugin contract); the update arithmetic runs in the
libflowse element-wise kernel.  When ``VF_fn`` is backed by the B200 NCSN++ (flowmse_b200 VFModel / NCSNpp),
``get_white_box_solver`` bypasses this per-step path and runs the whole loop inside ``flowse_sample``.
"""
import abc

import torch

from ..util.registry import Registry
from ..runtime import get_context

ODEsolverRegistry = Registry("ODEsolver")


def _as_float(v) -> float:
    return float(v.item()) if isinstance(v, torch.Tensor) else float(v)


class ODEsolver(abc.ABC):
    solver_id = None   # libflowse solver enum when the fused sampler supports this rule

    def __init__(self, ode, VF_fn):
        super().__init__()
        self.ode = ode
        self.VF_fn = VF_fn

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        ...


@ODEsolverRegistry.register("euler")
class EulerODEsolver(ODEsolver):
    solver_id = 0

    def update_fn(self, x, t, y, stepsize, *args):
        # dt = -stepsize; x + VF_fn(x,t,y) * dt   (odesolvers.py:42-47)
        v = self.VF_fn(x, t, y)
        return get_context(x.device).euler_step(x.contiguous(), v.contiguous(), _as_float(stepsize))


@ODEsolverRegistry.register("heun")
class HeunODEsolver(ODEsolver):
    """x_next = x + dt v0; x = x + dt/2 (v0 + VF(x_next, t+dt)).  Not in the reference source (stale .pyc only)."""
    solver_id = 1

    def update_fn(self, x, t, y, stepsize, *args):
        dt = -stepsize
        v0 = self.VF_fn(x, t, y)
        x_next = x + dt * v0
        return x + dt / 2 * (v0 + self.VF_fn(x_next, t + dt, y))


@ODEsolverRegistry.register("midpoint")
class MidpointODEsolver(ODEsolver):
    """x = x + dt VF(x + dt/2 VF(x,t), t + dt/2).  Not in the reference source (stale .pyc only)."""
    solver_id = 2

    def update_fn(self, x, t, y, stepsize, *args):
        dt = -stepsize
        x_mid = x + dt / 2 * self.VF_fn(x, t, y)
        return x + dt * self.VF_fn(x_mid, t + dt / 2, y)
