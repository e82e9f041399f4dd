"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from /root/reference).

Run here (CPU container) with:  python oracle/gen_golden.py
The reference is a Python package, so it cannot travel to the GPU box; the vectors
written by this script are committed instead.  Nothing in tests/, bench.py or smoke()
reads /root/reference at run time.

Stubs: pytorch_lightning, torch_ema, matplotlib, pesq, pystoi are absent from this image
and are not on the arithmetic path; torch.utils.cpp_extension.load is
stubbed so that importing ncsnpp_utils/op does not JIT-compile the CUDA extensions (the
CPU path uses upfirdn2d_native, op/upfirdn2d.py:146-149).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)


def install_stubs():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl

    te = types.ModuleType("torch_ema")

    class ExponentialMovingAverage:
        """Minimal torch_ema 0.3 stand-in: shadow copy of the parameters."""

        def __init__(self, parameters, decay):
            self.decay = decay
            self.shadow_params = [p.clone().detach() for p in parameters]
            self.collected_params = None

        def store(self, parameters):
            self.collected_params = [p.clone() for p in parameters]

        def copy_to(self, parameters):
            for s, p in zip(self.shadow_params, parameters):
                p.data.copy_(s.data)

        def restore(self, parameters):
            for c, p in zip(self.collected_params, parameters):
                p.data.copy_(c.data)

        def to(self, *a, **k):
            pass

        def state_dict(self):
            return dict(decay=self.decay, num_updates=0, shadow_params=self.shadow_params,
                        collected_params=self.collected_params)

        def load_state_dict(self, sd):
            self.shadow_params = [p.clone() for p in sd["shadow_params"]]

        def update(self, parameters):
            pass

    te.ExponentialMovingAverage = ExponentialMovingAverage
    sys.modules["torch_ema"] = te

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    pesq = types.ModuleType("pesq")
    pesq.pesq = lambda *a, **k: float("nan")
    sys.modules["pesq"] = pesq
    pystoi = types.ModuleType("pystoi")
    pystoi.stoi = lambda *a, **k: float("nan")
    sys.modules["pystoi"] = pystoi

    import torch.utils.cpp_extension as cpp
    cpp.load = lambda *a, **k: types.SimpleNamespace()


def hexf(v: float) -> str:
    return float(np.float32(v)).hex()


def tap_summary(t: torch.Tensor) -> np.ndarray:
    """[mean, mean|.|, max|.|] + a strided sample, for layer-wise parity."""
    t = t.detach().float()
    stats = torch.tensor([t.mean(), t.abs().mean(), t.abs().max()])
    samp = t[:, ::17, ::5, ::7].reshape(-1)[:512]
    return torch.cat([stats, samp]).numpy()


def main():
    install_stubs()
    sys.path.insert(0, REF)
    torch.set_grad_enabled(False)
    torch.manual_seed(0)

    from flowmse.backbones.ncsnpp import NCSNpp
    from flowmse.backbones.ncsnpp_utils import up_or_down_sampling as uds
    from flowmse.odes import FLOWMATCHING
    from flowmse.sampling import get_white_box_solver
    from flowmse.util.other import pad_spec
    from flowmse.model import VFModel
    from flowmse.data_module import SpecsDataModule

    from flowmse_b200.checkpoint import synthetic_state_dict
    from flowmse_b200 import ncsnpp_spec
    from oracle import ncsnpp_oracle as orc

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---- 1. state-dict layout of the real reference module -------------------------------------
    ref_net = NCSNpp()
    ref_keys = [(k, tuple(v.shape)) for k, v in ref_net.state_dict().items()]
    ours = ncsnpp_spec.state_dict_layout()
    assert ref_keys == [(k, tuple(s)) for k, s in ours], "state_dict layout mismatch vs reference"
    print("layout OK:", len(ref_keys), "tensors,", sum(v.numel() for v in ref_net.state_dict().values()), "params")
    with open(os.path.join(out_dir, "state_dict_layout.txt"), "w") as f:
        for k, s in ref_keys:
            f.write(f"{k} {'x'.join(map(str, s))}\n")

    sd = synthetic_state_dict(0)
    ref_net.load_state_dict(sd, strict=True)
    ref_net.eval()

    # ---- 2. schedule ------------------------------------------------------------------------------
    sched = {}
    for N in (1, 2, 5, 25, 30):
        ts = torch.linspace(1.0, 0.03, N)
        steps = [(ts[i] - ts[i + 1]) if i != N - 1 else ts[-1] for i in range(N)]
        sched[f"t{N}"] = ts.numpy()
        sched[f"s{N}"] = torch.stack(steps).numpy()
    np.savez(os.path.join(out_dir, "schedule.npz"), **sched)
    print("N=5 timesteps:", [hexf(v) for v in sched["t5"]])

    # ---- 3. FIR resampling against the reference's upfirdn2d_native -------------------------------
    g = torch.Generator().manual_seed(11)
    xf = torch.randn(2, 3, 8, 12, generator=g)
    up_ref = uds.upsample_2d(xf, (1, 3, 3, 1), factor=2)
    dn_ref = uds.downsample_2d(xf, (1, 3, 3, 1), factor=2)
    print("FIR up   oracle-vs-ref max err:", (orc.fir_upsample2(xf) - up_ref).abs().max().item())
    print("FIR down oracle-vs-ref max err:", (orc.fir_downsample2(xf) - dn_ref).abs().max().item())
    np.savez(os.path.join(out_dir, "fir.npz"), x=xf.numpy(), up=up_ref.numpy(), down=dn_ref.numpy())

    # ---- 4. one NFE, T=64, B=2 (different t per batch element) -----------------------------------
    g = torch.Generator().manual_seed(21)
    B, Fq, T = 2, 256, 64
    xin = torch.view_as_complex(0.3 * torch.randn(B, 2, Fq, T, 2, generator=g))
    tt = torch.tensor([0.757, 0.03], dtype=torch.float32)

    taps_ref = {}
    hooks = []
    for i, mod in enumerate(ref_net.all_modules):
        kind = ncsnpp_spec.module_list()[i]["kind"]
        if kind in ("rb", "attn", "combine"):
            hooks.append(mod.register_forward_hook(lambda m, a, o, i=i: taps_ref.__setitem__(f"m{i}", o)))
    v_ref = ref_net(xin, tt)
    for h in hooks:
        h.remove()
    taps_orc = {}
    v_orc = orc.ncsnpp_forward(sd, xin, tt, taps=taps_orc)
    err = (v_orc - v_ref).abs().max().item()
    print(f"NFE  oracle-vs-ref max abs err: {err:.3e}   (|ref| mean {v_ref.abs().mean().item():.3f}, max {v_ref.abs().max().item():.3f})")
    worst = max(((taps_orc[k] - taps_ref[k]).abs().max().item(), k) for k in taps_ref)
    print("worst per-module oracle-vs-ref err:", worst)
    fw = dict(x=torch.view_as_real(xin).numpy(), t=tt.numpy(), v=torch.view_as_real(v_ref).numpy())
    for k, v in taps_ref.items():
        fw["tap_" + k] = tap_summary(v)
    np.savez_compressed(os.path.join(out_dir, "forward_T64.npz"), **fw)

    # ---- 5. full sampler through the reference's VFModel + get_white_box_solver ---------------------
    model = VFModel(backbone="ncsnpp", ode="flowmatching", data_module_cls=SpecsDataModule, base_dir="")
    model.dnn.load_state_dict(sd, strict=True)
    model.ema.shadow_params = [p.clone().detach() for p in model.parameters()]   # pitfall 6
    model.eval()
    ode = model.ode
    assert isinstance(ode, FLOWMATCHING)
    g = torch.Generator().manual_seed(31)
    Y = torch.view_as_complex(0.3 * torch.randn(1, 1, 256, 64, 2, generator=g))
    samp = dict(Y=torch.view_as_real(Y).numpy())
    for N in (1, 5):
        torch.manual_seed(1234)
        z_expected = torch.randn_like(Y)
        torch.manual_seed(1234)
        sampler = get_white_box_solver("euler", ode, model, Y=Y, Y_prior=Y, T_rev=1.0, t_eps=0.03, N=N)
        x_ref, ns = sampler()
        assert ns == N
        x_orc = orc.sample(sd, Y, z_expected, N)
        print(f"sampler N={N} oracle-vs-ref max abs err: {(x_orc - x_ref).abs().max().item():.3e}")
        samp[f"x_euler_N{N}"] = torch.view_as_real(x_ref).numpy()
        samp["z"] = torch.view_as_real(z_expected).numpy()
    # Heun / midpoint: restated rule driven by the reference's own VFModel.forward
    for solver, N in (("heun", 3), ("midpoint", 3)):
        z = torch.view_as_complex(torch.from_numpy(samp["z"]))
        x_h = orc.sample(sd, Y, z, N, solver=solver, vf=lambda x, t, y: model(x, t, y))
        x_o = orc.sample(sd, Y, z, N, solver=solver)
        print(f"{solver} N={N} oracle-vf vs reference-vf max abs err: {(x_h - x_o).abs().max().item():.3e}")
        samp[f"x_{solver}_N{N}"] = torch.view_as_real(x_h).numpy()
    np.savez_compressed(os.path.join(out_dir, "sampler_T64.npz"), **samp)

    # ---- 6. plumbing config 1: 4 s synthetic wav -> STFT -> transform -> pad_spec shape -------------
    torch.manual_seed(0)
    n = 64000
    tgrid = torch.arange(n) / 16000.0
    wav = 0.1 * torch.randn(1, n) + 0.5 * torch.sin(2 * np.pi * 220 * tgrid) + 0.25 * torch.sin(2 * np.pi * 1320 * tgrid)
    wav = wav / wav.abs().max()
    Yw = torch.unsqueeze(model._forward_transform(model._stft(wav)), 0)
    Yp = pad_spec(Yw)
    print("config-1 spectrogram:", tuple(Yw.shape), "->", tuple(Yp.shape))
    np.savez_compressed(os.path.join(out_dir, "stft_cfg1.npz"), wav=wav.numpy()[:, :4096],
                        shape=np.array(Yw.shape), padded_shape=np.array(Yp.shape),
                        Y_head=torch.view_as_real(Yp[0, 0, :, :8]).numpy())
    print("done ->", out_dir)


if __name__ == "__main__":
    main()
