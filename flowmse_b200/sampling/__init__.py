"""Drop-in for /root/reference/flowmse/sampling/__init__.py: get_white_box_solver (same signature and return).

Fast path: when ``VF_fn`` is a flowmse_b200 ``VFModel`` / ``NCSNpp`` (i.e. the vector field is the B200 backbone)
the prior sample, the N solver steps and all network evaluations run inside one C-ABI call (``flowse_sample``):
no Python between steps, each NFE replayed as a CUDA graph.  Otherwise the reference's Python loop is reproduced
step by step with the registered ``ODEsolver`` (any callable ``VF_fn``), the update arithmetic still on the GPU.
"""
import torch

from .odesolvers import ODEsolver, ODEsolverRegistry

__all__ = ["ODEsolverRegistry", "ODEsolver", "get_white_box_solver", "get_black_box_solver", "get_pc_sampler",
           "timesteps_and_stepsizes"]


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


# ---------------------------------------------------------------------------------------------------------------
# Black-box adaptive solver (reference: get_black_box_solver, sampling/__init__.py:64-114)
# ---------------------------------------------------------------------------------------------------------------
# Dormand-Prince 5(4) tableau and step-size controller exactly as scipy.integrate.solve_ivp(method="RK45") applies them
# (scipy/integrate/_ivp/rk.py: RK45, rk_step, RungeKutta._step_impl; common.py: select_initial_step, norm).
_RK45_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
_RK45_A = ((), (1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9), (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
           (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
_RK45_B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
_RK45_E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
_SAFETY, _MIN_FACTOR, _MAX_FACTOR, _ERR_EXP = 0.9, 0.2, 10.0, -1.0 / 5.0


def _rk45_on_device(VF_fn, x0, y, T_rev, t_eps, rtol, atol):
    """solve_ivp(ode_func, (T_rev, t_eps), x0, method="RK45", rtol, atol) with the state on the GPU.

    The reference moves the whole state to the host and back for EVERY network evaluation and does the Runge-Kutta
    arithmetic in NumPy (complex128 state, complex64 derivatives).  Here the state (complex128) and the seven stage
    derivatives (complex64) stay in HBM, the stage combinations / the scaled error norm are one libflowse kernel each
    (flowse_rk_lincomb), and the only host round trip is the 8-byte error norm the step-size controller needs per step.
    Returns (y at t_eps as complex128, number of VF evaluations)."""
    import math
    import numpy as np
    from ..runtime import get_context
    ctx = get_context(x0.device)
    B = y.shape[0]
    n = x0.numel()
    nfev = 0

    def fun(t, x32):
        nonlocal nfev
        nfev += 1
        vec_t = torch.ones(B, device=x0.device) * t
        return VF_fn(x32, vec_t, y)

    state = torch.view_as_complex(torch.view_as_real(x0).double().contiguous())        # complex128, like solve_ivp's y
    new_state = torch.empty_like(state)
    K = torch.empty((7,) + tuple(x0.shape), dtype=torch.complex64, device=x0.device)   # stage derivatives
    xin = torch.empty_like(x0)
    t, t_bound = float(T_rev), float(t_eps)
    direction = float(np.sign(t_bound - t)) if t_bound != t else 1.0
    rms = lambda sumsq: math.sqrt(sumsq / n)
    K[0].copy_(fun(t, x0.contiguous()))
    # ---- select_initial_step (order = error estimator order = 4)
    interval = abs(t_bound - t)
    if interval == 0.0:
        return state, nfev
    d0 = rms(ctx.rk_lincomb(state, K, (), norm_of=(state, state), rtol=rtol, atol=atol))
    d1 = rms(ctx.rk_lincomb(None, K, (1.0,), norm_of=(state, state), rtol=rtol, atol=atol))
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval)
    ctx.rk_lincomb(state, K, (h0 * direction,), out32=xin)
    K[1].copy_(fun(t + h0 * direction, xin))
    d2 = rms(ctx.rk_lincomb(None, K, (-1.0, 1.0), norm_of=(state, state), rtol=rtol, atol=atol)) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1 / 5)
    h_abs = min(100 * h0, h1, interval)
    # ---- adaptive steps
    while direction * (t - t_bound) < 0:
        min_step = 10 * abs(float(np.nextafter(t, direction * np.inf)) - t)
        h_abs = max(h_abs, min_step)
        rejected = False
        while True:
            if h_abs < min_step:
                return state, nfev                   # scipy: TOO_SMALL_STEP, the last accepted state is returned
            h = h_abs * direction
            t_new = t + h
            if direction * (t_new - t_bound) > 0:
                t_new = t_bound
            h = t_new - t
            h_abs = abs(h)
            for s in range(1, 6):                    # stages: K[s] = f(t + c_s h, y + h sum_j a_sj K[j])
                ctx.rk_lincomb(state, K, [a * h for a in _RK45_A[s]], out32=xin)
                K[s].copy_(fun(t + _RK45_C[s] * h, xin))
            ctx.rk_lincomb(state, K, [b * h for b in _RK45_B], out64=new_state, out32=xin)
            K[6].copy_(fun(t + h, xin))              # f_new (first stage of the next step)
            err = rms(ctx.rk_lincomb(None, K, [e * h for e in _RK45_E], norm_of=(state, new_state), rtol=rtol, atol=atol))
            if err < 1:
                factor = _MAX_FACTOR if err == 0 else min(_MAX_FACTOR, _SAFETY * err ** _ERR_EXP)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                break
            h_abs *= max(_MIN_FACTOR, _SAFETY * err ** _ERR_EXP)
            rejected = True
        t = t_new
        state, new_state = new_state, state
        K[0].copy_(K[6])
    return state, nfev


def get_black_box_solver(ode, VF_fn, y, rtol=1e-5, atol=1e-5, T_rev=1.0, t_eps=0.03, N=30, method="RK45", device="cuda",
                         **kwargs):
    """Probability-flow ODE sampler with a black-box adaptive solver: same signature and return ``(x, nfe)`` as the
    reference (sampling/__init__.py:64-114).  ``method="RK45"`` (the default) runs on the device (``_rk45_on_device``);
    any other ``scipy.integrate.solve_ivp`` method, or extra solve_ivp keyword arguments, take the reference's own route
    through scipy with one host round trip per evaluation."""

    def ode_solver(**_unused):
        with torch.no_grad():
            x = ode.prior_sampling(y.shape, y)[0].to(device)
            if method == "RK45" and not kwargs and x.is_cuda:
                state, nfe = _rk45_on_device(VF_fn, x, y.to(x.device), T_rev, t_eps, rtol, atol)
                return state.reshape(y.shape).to(device).type(torch.complex64), nfe
            from scipy import integrate

            def ode_func(t, flat):
                xt = torch.from_numpy(flat.reshape(tuple(y.shape))).to(device).type(torch.complex64)
                vec_t = torch.ones(y.shape[0], device=xt.device) * t
                return VF_fn(xt, vec_t, y).detach().cpu().numpy().reshape((-1,))

            sol = integrate.solve_ivp(ode_func, (T_rev, t_eps), x.detach().cpu().numpy().reshape((-1,)), rtol=rtol, atol=atol,
                                      method=method, **kwargs)
            out = torch.tensor(sol.y[:, -1]).reshape(y.shape).to(device).type(torch.complex64)
            return out, sol.nfev

    return ode_solver
