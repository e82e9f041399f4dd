"""Inference-side mirror of the reference's VFModel (/root/reference/flowmse/model.py:19-206).

Keeps what evaluate.py touches (evaluate.py:64-73,85-86,113,118,132): ``load_from_checkpoint`` reading the
Lightning + torch_ema layout (EMA weights become live, model.py:81-106), ``eval(no_ema=...)``, ``forward(x, t, y)``,
``.ode`` with ``sigma_min/sigma_max/T_rev``, ``.t_eps``, and the STFT helpers ``_stft/_istft/_forward_transform/
_backward_transform/to_audio`` (data_module.py:149-175,199-205).  Training (Lightning hooks, loss, optimiser, data
loaders) is out of scope.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch
import torch.nn as nn

from .backbones import BackboneRegistry
from .odes import ODERegistry
from . import checkpoint as ckpt_io


class SpecTransform:
    """STFT / iSTFT and the amplitude compression of SpecsDataModule (data_module.py:149-175,177-205)."""

    def __init__(self, n_fft=510, hop_length=128, spec_factor=0.15, spec_abs_exponent=0.5, transform_type="exponent",
                 window="hann", **ignored):
        if window not in ("hann", "sqrthann"):                 # get_window, data_module.py:13-19
            raise NotImplementedError(f"Window type {window} not implemented!")
        self.n_fft, self.hop_length, self.window = n_fft, hop_length, window
        self.spec_factor, self.spec_abs_exponent, self.transform_type = spec_factor, spec_abs_exponent, transform_type
        self._windows = {}

    def hparams(self):
        """The constructor arguments, e.g. to rebuild the same transform on another rank."""
        return dict(n_fft=self.n_fft, hop_length=self.hop_length, spec_factor=self.spec_factor,
                    spec_abs_exponent=self.spec_abs_exponent, transform_type=self.transform_type, window=self.window)

    def _window(self, x):
        w = self._windows.get(x.device)
        if w is None:
            w = torch.hann_window(self.n_fft, periodic=True)
            if self.window == "sqrthann":
                w = torch.sqrt(w)
            w = w.to(x.device)
            self._windows[x.device] = w
        return w

    def stft(self, sig):
        return torch.stft(sig, n_fft=self.n_fft, hop_length=self.hop_length, window=self._window(sig), center=True,
                          return_complex=True)

    def istft(self, spec, length=None):
        return torch.istft(spec, n_fft=self.n_fft, hop_length=self.hop_length, window=self._window(spec), center=True,
                           length=length)

    def spec_fwd(self, spec):
        if self.transform_type == "exponent":
            if self.spec_abs_exponent != 1:
                e = self.spec_abs_exponent
                spec = spec.abs() ** e * torch.exp(1j * spec.angle())
            spec = spec * self.spec_factor
        elif self.transform_type == "log":
            spec = torch.log(1 + spec.abs()) * torch.exp(1j * spec.angle())
            spec = spec * self.spec_factor
        return spec

    def spec_back(self, spec):
        if self.transform_type == "exponent":
            spec = spec / self.spec_factor
            if self.spec_abs_exponent != 1:
                e = self.spec_abs_exponent
                spec = spec.abs() ** (1 / e) * torch.exp(1j * spec.angle())
        elif self.transform_type == "log":
            spec = spec / self.spec_factor
            spec = (torch.exp(spec.abs()) - 1) * torch.exp(1j * spec.angle())
        return spec


@contextlib.contextmanager
def reference_pickle_shims():
    """Lightning checkpoints written by the reference pickle ``hyper_parameters['data_module_cls']`` BY REFERENCE as
    ``flowmse.data_module.SpecsDataModule`` (VFModel.__init__ keeps the class, model.py:34-56, and
    ``save_hyperparameters`` stores it), so ``torch.load`` must be able to import that name.  Where the reference package
    (or one of its training-only dependencies: pytorch_lightning, torch_ema) is not importable, placeholder modules are
    installed for the duration of the load; the class is never instantiated here (SpecTransform replaces it)."""
    created = []
    try:
        import flowmse.data_module  # noqa: F401  (the real package is importable: nothing to do)
    except Exception:
        pkg = sys.modules.get("flowmse")
        if pkg is None:
            pkg = types.ModuleType("flowmse")
            pkg.__path__ = []
            sys.modules["flowmse"] = pkg
            created.append("flowmse")
        if "flowmse.data_module" not in sys.modules:
            dm = types.ModuleType("flowmse.data_module")

            class SpecsDataModule:       # placeholder for unpickling only
                pass

            SpecsDataModule.__module__ = "flowmse.data_module"
            SpecsDataModule.__qualname__ = "SpecsDataModule"
            dm.SpecsDataModule = SpecsDataModule
            sys.modules["flowmse.data_module"] = dm
            pkg.data_module = dm
            created.append("flowmse.data_module")
    try:
        yield
    finally:
        for name in created:
            sys.modules.pop(name, None)


class VFModel(nn.Module):
    _flowse_fused = True

    def __init__(self, backbone="ncsnpp", ode="flowmatching", t_eps=0.03, T_rev=1.0, **kwargs):
        super().__init__()
        self.dnn = BackboneRegistry.get_by_name(backbone)(**kwargs)
        self.ode = ODERegistry.get_by_name(ode)(**kwargs)
        self.t_eps = t_eps
        self.T_rev = T_rev
        self.ode.T_rev = T_rev
        self.data_module = SpecTransform(**kwargs)
        self._ema_state = None       # shadow weights from the checkpoint (names -> tensors)
        self._live_backup = None

    # ---- checkpoint ------------------------------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", **override):
        """Read a reference checkpoint (Lightning 1.6.5 + torch_ema 0.3 layout, model.py:81-90)."""
        try:
            with reference_pickle_shims():
                ck = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        except Exception as e:
            raise RuntimeError(f"cannot unpickle {checkpoint_path}: {e}") from e
        return cls.from_checkpoint_dict(ck, **override)

    @classmethod
    def from_checkpoint_dict(cls, ck: dict, **override):
        hp = dict(ck.get("hyper_parameters", {}))
        hp.pop("data_module_cls", None)
        hp.update({k: v for k, v in override.items() if k not in ("base_dir", "batch_size", "num_workers", "kwargs")})
        allowed = {k: hp[k] for k in ("backbone", "ode", "t_eps", "T_rev", "sigma_min", "sigma_max", "n_fft",
                                      "hop_length", "spec_factor", "spec_abs_exponent", "transform_type", "window")
                   if k in hp}
        model = cls(**allowed)
        live = ckpt_io.backbone_state_from_checkpoint(ck, use_ema=False)
        model.dnn.load_state_dict(live, strict=True)
        if ck.get("ema") is not None:
            model._ema_state = ckpt_io.backbone_state_from_checkpoint(ck, use_ema=True)
        return model

    @classmethod
    def load_from_packed(cls, path, **override):
        """Build the model from a packed deployment checkpoint (``checkpoint.save_packed``): hyper-parameters as in
        ``from_checkpoint_dict``, weights handed to libflowse in its own fp16 hi/lo layout (no fp32 tensors, no host-side
        packing pass).  The result is for inference only."""
        ck = ckpt_io.load_packed(path)
        hp = dict(ck["hyper_parameters"])
        hp.update({k: v for k, v in override.items() if k not in ("base_dir", "batch_size", "num_workers", "kwargs")})
        allowed = {k: hp[k] for k in ("backbone", "ode", "t_eps", "T_rev", "sigma_min", "sigma_max", "n_fft",
                                      "hop_length", "spec_factor", "spec_abs_exponent", "transform_type", "window")
                   if k in hp}
        model = cls(**allowed)
        model.dnn.load_packed_weights(ck["blob"])
        return model.eval()

    def save_packed(self, path, device="cuda"):
        """Write the LIVE weights (the EMA weights after ``eval()``) as a packed deployment checkpoint."""
        dm = self.data_module
        hp = dict(backbone="ncsnpp", ode="flowmatching", t_eps=self.t_eps, T_rev=self.T_rev, sigma_min=self.ode.sigma_min,
                  sigma_max=self.ode.sigma_max, **dm.hparams())
        ckpt_io.save_packed(path, self.flowse_context(device), hp)

    # ---- EMA swap, as VFModel.train/eval (model.py:92-106) --------------------------------------------------------
    def train(self, mode=True, no_ema=False):
        res = super().train(mode)
        if self._ema_state is not None:
            if mode is False and not no_ema:
                if self._live_backup is None:
                    self._live_backup = {k: v.detach().clone() for k, v in self.dnn.state_dict().items()}
                self.dnn.load_state_dict(self._ema_state, strict=True)
            elif self._live_backup is not None:
                self.dnn.load_state_dict(self._live_backup, strict=True)
                self._live_backup = None
        return res

    def eval(self, no_ema=False):
        return self.train(False, no_ema=no_ema)

    # ---- hot path ----------------------------------------------------------------------------------------------
    def flowse_context(self, device):
        return self.dnn.flowse_context(device)

    def forward(self, x, t, y):
        """-dnn(cat([x, y], 1), t) (model.py:164-170) without materialising the concat."""
        with torch.no_grad():
            return self.flowse_context(x.device).vf_forward(x.contiguous(), t, y.contiguous())

    def enhance_spec(self, Y, N=5, odesolver="euler", T_rev=None, t_eps=None):
        """Run the reverse-ODE sampler on a padded spectrogram batch [B,1,256,T]; returns the enhanced spectrogram."""
        from .sampling import get_white_box_solver
        sampler = get_white_box_solver(odesolver, self.ode, self, Y=Y, Y_prior=Y,
                                       T_rev=self.T_rev if T_rev is None else T_rev,
                                       t_eps=self.t_eps if t_eps is None else t_eps, N=N)
        return sampler()[0]

    def enhance(self, y, N=5, odesolver="euler", **kw):
        """Waveform in, waveform out: evaluate.py:107-136 for one utterance ([1, samples] tensor on a CUDA device)."""
        from .util.other import pad_spec
        T_orig = y.size(1)
        norm = y.abs().max()
        Y = torch.unsqueeze(self._forward_transform(self._stft(y / norm)), 0)
        Y = pad_spec(Y)
        sample = self.enhance_spec(Y.contiguous(), N=N, odesolver=odesolver, **kw)
        return self.to_audio(sample.squeeze(), T_orig) * norm

    def _device_stft_ok(self):
        dm = self.data_module
        # the device STFT kernels are specialised to the frame geometry the backbone needs (n_fft 510 -> F = 256, hop
        # 128); every transform_type and window of the reference is handled on the device
        return (dm.n_fft == 510 and dm.hop_length == 128 and dm.transform_type in ("exponent", "log", "none")
                and dm.window in ("hann", "sqrthann"))

    def enhance_batch(self, wavs, N=5, odesolver="euler", normalize=True, **kw):
        """evaluate.py:97-136 for a list of 1-D waveforms of ANY lengths on one CUDA device, without the per-file host
        round trips: ONE batched device STFT (+ peak normalisation, amplitude compression, pad_spec) for all of them
        (flowse_stft_spec), one sampler call per frame-count bucket (utterances whose padded T agree share a batch), one
        batched device iSTFT (flowse_spec_istft).  Returns a list of enhanced waveforms, input order and lengths kept.
        The prior noise is drawn per bucket in bucket order from torch's global generator of the device."""
        if not wavs:
            return []
        if not self._device_stft_ok():
            return [self.enhance(w.reshape(1, -1), N=N, odesolver=odesolver, **kw).reshape(-1) for w in wavs]
        dev = wavs[0].device
        ctx = self.flowse_context(dev)
        dm = self.data_module
        lens = [int(w.numel()) for w in wavs]
        # bucket by padded frame count (pad_spec: multiples of 64), longest bucket first
        buckets = {}
        for i, L in enumerate(lens):
            buckets.setdefault(((ctx.frames_of(L) + 63) // 64) * 64, []).append(i)
        out = [None] * len(wavs)
        for Tpad in sorted(buckets, reverse=True):
            idx = buckets[Tpad]
            Lb = [lens[i] for i in idx]
            wav = torch.zeros((len(idx), max(Lb)), dtype=torch.float32, device=dev)
            for r, i in enumerate(idx):
                wav[r, :lens[i]] = wavs[i].reshape(-1).to(device=dev, dtype=torch.float32)
            Y, peak = ctx.stft_spec(wav, Lb, normalize=normalize, spec_factor=dm.spec_factor,
                                    abs_exponent=dm.spec_abs_exponent, Tpad=Tpad, transform_type=dm.transform_type,
                                    window=dm.window)
            X = self.enhance_spec(Y, N=N, odesolver=odesolver, **kw)
            x_hat = ctx.spec_istft(X.contiguous(), Lb, peak=peak if normalize else None, spec_factor=dm.spec_factor,
                                   abs_exponent=dm.spec_abs_exponent, transform_type=dm.transform_type,
                                   window=dm.window)
            for r, i in enumerate(idx):
                out[i] = x_hat[r, :lens[i]]
        return out

    # ---- STFT helpers (model.py:190-203) -----------------------------------------------------------------------
    def to_audio(self, spec, length=None):
        return self._istft(self._backward_transform(spec), length)

    def _forward_transform(self, spec):
        return self.data_module.spec_fwd(spec)

    def _backward_transform(self, spec):
        return self.data_module.spec_back(spec)

    def _stft(self, sig):
        return self.data_module.stft(sig)

    def _istft(self, spec, length=None):
        return self.data_module.istft(spec, length)


ScoreModel = VFModel   # name used by BASELINE.json's north_star (SGMSE lineage)
