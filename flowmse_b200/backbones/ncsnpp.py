"""B200-native NCSN++ vector-field backbone behind the reference's plugin interface.

Mirror of /root/reference/flowmse/backbones/ncsnpp.py: registered as 'ncsnpp' in ``BackboneRegistry``, constructed
as ``NCSNpp(**kwargs)`` (unknown kwargs ignored, ncsnpp.py:66), ``forward(x [B,2,F,T] complex64, time_cond [B])``
-> ``[B,1,F,T] complex64``, static ``add_argparse_args``.  The module tree only HOLDS parameters, under exactly the
reference's names (``all_modules.<i>.<...>``, ``output_layer.*``), so reference checkpoints load with
``load_state_dict``; the arithmetic is one C-ABI call into libflowse (hand-written sm_100a kernels).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .shared import BackboneRegistry
from .. import ncsnpp_spec as spec
from ..lib import FlowseError
from ..runtime import new_context


class _ParamHolder(nn.Module):
    """A module that owns parameters under dotted sub-names ('Conv_0.weight' -> self.Conv_0.weight)."""

    def __init__(self, params):
        super().__init__()
        for sub, shape in params:
            parts = sub.split(".")
            mod = self
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, nn.Module())
                mod = getattr(mod, p)
            mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))

    def forward(self, *a, **k):
        raise FlowseError("NCSN++ sub-modules only hold parameters; call NCSNpp.forward")


@BackboneRegistry.register("ncsnpp")
class NCSNpp(nn.Module):
    _flowse_fused = True    # get_white_box_solver may run the whole sampler in libflowse with this vector field

    @staticmethod
    def add_argparse_args(parser):
        return parser

    def __init__(self, scale_by_sigma=True, nonlinearity="swish", nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2),
                 num_res_blocks=2, attn_resolutions=(16,), resamp_with_conv=True, conditional=True, fir=True,
                 fir_kernel="song", skip_rescale=True, resblock_type="biggan", progressive="output_skip",
                 progressive_input="input_skip", progressive_combine="sum", init_scale=0.0, fourier_scale=16,
                 image_size=256, embedding_type="fourier", dropout=0.0, **unused_kwargs):
        super().__init__()
        cfg = dict(nonlinearity=nonlinearity, nf=nf, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks,
                   attn_resolutions=tuple(attn_resolutions), conditional=conditional, fir=fir,
                   skip_rescale=skip_rescale, resblock_type=resblock_type.lower(), progressive=progressive.lower(),
                   progressive_input=progressive_input.lower(), progressive_combine=progressive_combine.lower(),
                   image_size=image_size, embedding_type=embedding_type.lower(), dropout=float(dropout))
        want = dict(nonlinearity="swish", nf=spec.NF, ch_mult=spec.CH_MULT, num_res_blocks=spec.NUM_RES_BLOCKS,
                    attn_resolutions=spec.ATTN_RESOLUTIONS, conditional=True, fir=True, skip_rescale=True,
                    resblock_type="biggan", progressive="output_skip", progressive_input="input_skip",
                    progressive_combine="sum", image_size=spec.IMAGE_SIZE, embedding_type="fourier", dropout=0.0)
        bad = {k: v for k, v in cfg.items() if want[k] != v}
        if bad:
            raise NotImplementedError(f"the B200 NCSN++ kernels are specialised to the FlowSE default config; got {bad}")
        self.output_layer = _ParamHolder([("weight", (2, spec.NUM_CHANNELS, 1, 1)), ("bias", (2,))])
        self.all_modules = nn.ModuleList([_ParamHolder(spec.module_params(m)) for m in spec.module_list()])
        self._ctx = {}           # device index -> (Context, version stamp)
        self._stamp, self._sentinels = 0, None
        self._packed_blob = None   # libflowse packed weights (checkpoint.load_packed): used instead of the fp32 parameters
        self._reset_like_reference(fourier_scale)

    def _reset_like_reference(self, fourier_scale):
        """Cheap stand-in for the reference's initialisers (layers.py:54-91): DDPM fan-avg uniform for weights,
        zeros for biases, ones for GroupNorm scales, N(0, scale^2) Fourier frequencies.  Real use loads a checkpoint."""
        with torch.no_grad():
            for name, p in self.named_parameters():
                leaf = name.rsplit(".", 1)[-1]
                if name == "all_modules.0.W":
                    p.copy_(torch.randn(p.shape) * fourier_scale)
                elif p.dim() == 1 and leaf == "weight":
                    p.fill_(1.0)
                elif leaf in ("bias", "b"):
                    p.zero_()
                else:
                    shape = p.shape
                    rf = 1
                    for s in shape[2:]:
                        rf *= s
                    fan_in, fan_out = (shape[1] * rf, shape[0] * rf) if leaf == "weight" else (shape[0], shape[1])
                    bound = (3.0 / ((fan_in + fan_out) / 2.0)) ** 0.5
                    p.uniform_(-bound, bound)

    # ---- libflowse context management --------------------------------------------------------------------------
    def _version(self):
        """Stamp of the live weights.  Every bulk update path (load_state_dict, an EMA swap, optimiser steps) rewrites ALL
        parameters in place, which bumps each tensor's autograd version counter; a few sentinels therefore identify the
        state without walking all 647 tensors on every sampler call.  ``invalidate()`` forces a re-pack after any other
        kind of edit."""
        s = self._sentinels
        if s is None:
            ps = list(self.parameters())
            s = self._sentinels = [ps[0], ps[len(ps) // 3], ps[(2 * len(ps)) // 3], ps[-1]]
        return (self._stamp,) + tuple(p._version for p in s) + tuple(p.data_ptr() for p in s)

    def invalidate(self):
        """Declare the parameters changed: the next call re-packs them into the libflowse context."""
        self._stamp += 1
        self._sentinels = None

    def load_packed_weights(self, blob):
        """Use a libflowse packed-weights blob (``Context.export_packed``) as this backbone's weights.  The fp32
        ``nn.Parameter`` tensors are NOT updated (the packed format is the deployment format; keep the fp32 checkpoint for
        anything that needs ``state_dict``); a later ``load_state_dict`` switches back to the fp32 parameters."""
        self._packed_blob = blob
        self._stamp += 1
        self._sentinels = None

    def _load_from_state_dict(self, *a, **k):
        self._packed_blob = None
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .half() replace the parameter tensors
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def flowse_context(self, device):
        """The libflowse context holding THIS module's weights on `device` (packed on first use, re-packed when a
        parameter changed, e.g. after load_state_dict or an EMA swap)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise FlowseError("the B200 NCSN++ runs on CUDA tensors only; there is no CPU fallback")
        idx = torch.cuda.current_device() if dev.index is None else dev.index
        ver = self._version()
        ent = self._ctx.get(idx)
        if ent is None or ent[1] != ver:
            if ent is not None:
                ent[0].close()
            ctx = new_context(dev)
            if self._packed_blob is not None:
                ctx.load_packed(self._packed_blob)
            else:
                ctx.load_state_dict({k: v for k, v in self.state_dict().items()})
            self._ctx[idx] = (ctx, ver)
            ent = self._ctx[idx]
        return ent[0]

    def forward(self, x, time_cond):
        """x: complex64 [B,2,256,T] = cat([x_t, y], 1); time_cond: [B].  Returns complex64 [B,1,256,T]."""
        with torch.no_grad():
            return self.flowse_context(x.device).ncsnpp_forward(x.contiguous(), time_cond)
