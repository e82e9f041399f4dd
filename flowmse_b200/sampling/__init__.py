"""Drop-in for /root/reference/flowmse/sampling/__init__.py: get_white_box_solver (same signature and return).

Fast path: when ``VF_fn`` is a flowmse_b200 ``VFModel`` / ``NCSNpp`` (i.e. the vector field is the B200 backbone)
the prior sample, the N solver steps and all network evaluations run inside one C-ABI call (``flowse_sample``):
no Python between steps, each NFE replayed as a CUDA graph.  Otherwise the reference's Python loop is reproduced
step by step with the registered ``ODEsolver`` (any callable ``VF_fn``), the update arithmetic still on the GPU.
"""
import torch

from .odesolvers import ODEsolver, ODEsolverRegistry

__all__ = ["ODEsolverRegistry", "ODEsolver", "get_white_box_solver", "get_pc_sampler", "timesteps_and_stepsizes"]


def timesteps_and_stepsizes(N, T_rev=1.0, t_eps=0.03, device="cpu"):
    """The reference schedule, bit for bit: torch.linspace(T_rev, t_eps, N) in fp32 and its fp32 differences;
    the last step size is t_{N-1} itself so the trajectory lands on t = 0 (sampling/__init__.py:45-53)."""
    timesteps = torch.linspace(T_rev, t_eps, N, device=device)
    steps = [timesteps[i] - timesteps[i + 1] if i != N - 1 else timesteps[-1] for i in range(N)]
    return timesteps, torch.stack(steps)


_schedule_cache = {}


def _host_schedule(T_rev, t_eps, N, device):
    """Host copy (tuple of fp32 values) of ``torch.linspace(T_rev, t_eps, N, device=device)`` - the tensor the reference
    builds on Y's device (sampling/__init__.py:45).  Reading it back is a device synchronisation, so it is done once per
    distinct (T_rev, t_eps, N, device type) and cached: repeated sampler() calls stay sync-free."""
    key = (float(T_rev), float(t_eps), int(N), torch.device(device).type)
    ts = _schedule_cache.get(key)
    if ts is None:
        ts = tuple(torch.linspace(T_rev, t_eps, N, device=device).cpu().tolist())
        _schedule_cache[key] = ts
    return ts


def _fused_backend(VF_fn):
    """The object owning a libflowse context if VF_fn is the B200 vector field, else None."""
    return VF_fn if getattr(VF_fn, "_flowse_fused", False) else None


def get_white_box_solver(odesolver_name, ode, VF_fn, Y, Y_prior=None, T_rev=1.0, t_eps=0.03, N=30, **kwargs):
    odesolver_cls = ODEsolverRegistry.get_by_name(odesolver_name)   # ValueError for unknown names, like the reference
    odesolver = odesolver_cls(ode, VF_fn)
    fused = _fused_backend(VF_fn)

    def ode_solver(Y_prior=Y_prior):
        with torch.no_grad():
            if Y_prior is None:
                Y_prior = Y
            if fused is not None and odesolver_cls.solver_id is not None and Y_prior.shape == Y.shape:
                z = torch.randn_like(Y_prior)                 # same generator call as ode.prior_sampling
                ts = _host_schedule(T_rev, t_eps, N, Y.device)
                x = fused.flowse_context(Y.device).sample(Y.contiguous(), z, ts,
                                                          solver=odesolver_cls.solver_id, sigma=ode.prior_std(),
                                                          y_prior=None if Y_prior is Y else Y_prior.contiguous())
                return x, len(ts)
            timesteps = torch.linspace(T_rev, t_eps, N, device=Y.device)
            xt, _ = ode.prior_sampling(Y_prior.shape, Y_prior)
            xt = xt.to(Y_prior.device)
            last_euler = ODEsolverRegistry.get_by_name("euler")(ode, VF_fn)
            for i in range(len(timesteps)):
                t = timesteps[i]
                last = i == len(timesteps) - 1
                stepsize = timesteps[-1] if last else t - timesteps[i + 1]
                vec_t = torch.ones(Y.shape[0], device=Y.device) * t
                # Heun / midpoint would evaluate the network at t = 0 on the last interval: Euler there
                step_fn = last_euler if (last and odesolver_name != "euler") else odesolver
                xt = step_fn.update_fn(xt, vec_t, Y, stepsize)
            return xt, len(timesteps)

    return ode_solver


def get_pc_sampler(*args, **kwargs):
    """Alias kept for callers written against the SGMSE ancestor's name (BASELINE.json north_star)."""
    return get_white_box_solver(*args, **kwargs)
