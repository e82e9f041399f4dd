"""ctypes binding of libflowse.so (include/flowse.h).  PyTorch tensors in, PyTorch tensors out.

There is no CPU path: every function here raises if the shared library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import ncsnpp_spec as spec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOWSE_LIB") or os.path.join(_HERE, "libflowse.so")   # FLOWSE_LIB: A/B builds of the library
_lib = None


class FlowseError(RuntimeError):
    pass


class _TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_longlong), ("numel", C.c_longlong)]


def load_library():
    """dlopen libflowse.so (built in-tree by flowmse_b200/csrc/build.sh) and declare its C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlowseError(f"{LIB_PATH} not found: build it with flowmse_b200/csrc/build.sh "
                          f"(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll, f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
    lib.flowse_create.argtypes = [C.POINTER(vp), i]; lib.flowse_create.restype = i
    lib.flowse_destroy.argtypes = [vp]; lib.flowse_destroy.restype = None
    lib.flowse_last_error.argtypes = [vp]; lib.flowse_last_error.restype = C.c_char_p
    lib.flowse_load_weights.argtypes = [vp, vp, C.POINTER(_TensorDesc), i]; lib.flowse_load_weights.restype = i
    lib.flowse_workspace_bytes.argtypes = [vp, i, i]; lib.flowse_workspace_bytes.restype = C.c_size_t
    lib.flowse_prior_sample.argtypes = [vp, vp, vp, f, vp, ll, vp]; lib.flowse_prior_sample.restype = i
    lib.flowse_ncsnpp_forward.argtypes = [vp, vp, ll, vp, ll, vp, vp, i, i, i, vp]; lib.flowse_ncsnpp_forward.restype = i
    lib.flowse_euler_step.argtypes = [vp, vp, vp, f, vp, ll, vp]; lib.flowse_euler_step.restype = i
    lib.flowse_sample.argtypes = [vp, vp, vp, vp, C.POINTER(f), i, i, f, vp, i, i, vp]; lib.flowse_sample.restype = i
    lib.flowse_set_option.argtypes = [vp, C.c_char_p, i]; lib.flowse_set_option.restype = i
    lib.flowse_kernel_launches.argtypes = [vp]; lib.flowse_kernel_launches.restype = ll
    lib.flowse_debug_tap.argtypes = [vp, i, C.POINTER(vp), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    lib.flowse_debug_tap.restype = i
    lib.flowse_profile_forward.argtypes = [vp, i, vp, vp, vp, vp, C.POINTER(i)]; lib.flowse_profile_forward.restype = i
    lib.flowse_debug_copy.argtypes = [vp, vp, vp, C.c_size_t]; lib.flowse_debug_copy.restype = i
    lib.flowse_pack_conv_weights.argtypes = [vp, i, i, i, vp, i, i, vp, C.POINTER(i)]
    lib.flowse_pack_conv_weights.restype = i
    lib.flowse_op_gn_prep.argtypes = [vp, vp, i, vp, i, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp]
    lib.flowse_op_gn_prep.restype = i
    lib.flowse_op_conv_gemm.argtypes = [vp, vp, i, i, vp, i, vp, i, i, vp, i, vp, i, vp, i, i, i, i, i, i, vp]
    lib.flowse_op_conv_gemm.restype = i
    lib.flowse_op_attention.argtypes = [vp, i, vp, vp, i, i, i, vp]; lib.flowse_op_attention.restype = i
    lib.flowse_op_head_conv.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, vp]; lib.flowse_op_head_conv.restype = i
    lib.flowse_fp16_overflow.argtypes = [vp, C.POINTER(ll), i]; lib.flowse_fp16_overflow.restype = i
    lib.flowse_rk_lincomb.argtypes = [vp, vp, vp, ll, C.POINTER(C.c_double), i, vp, vp, vp, vp, C.c_double, C.c_double,
                                      C.POINTER(C.c_double), ll, vp]
    lib.flowse_rk_lincomb.restype = i
    lib.flowse_packed_bytes.argtypes = [vp]; lib.flowse_packed_bytes.restype = C.c_size_t
    lib.flowse_export_packed.argtypes = [vp, vp, C.c_size_t]; lib.flowse_export_packed.restype = i
    lib.flowse_load_packed.argtypes = [vp, vp, C.c_size_t]; lib.flowse_load_packed.restype = i
    lib.flowse_stft_spec.argtypes = [vp, vp, ll, C.POINTER(i), i, i, f, f, vp, i, vp, vp]; lib.flowse_stft_spec.restype = i
    lib.flowse_spec_istft.argtypes = [vp, vp, i, C.POINTER(i), i, f, f, vp, vp, ll, vp]; lib.flowse_spec_istft.restype = i
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "flowse_create", "flowse_destroy", "flowse_last_error", "flowse_load_weights", "flowse_workspace_bytes",
    "flowse_prior_sample", "flowse_ncsnpp_forward", "flowse_euler_step", "flowse_sample", "flowse_set_option",
    "flowse_kernel_launches", "flowse_profile_forward", "flowse_debug_tap", "flowse_debug_copy", "flowse_pack_conv_weights", "flowse_op_gn_prep",
    "flowse_op_conv_gemm", "flowse_op_attention", "flowse_stft_spec", "flowse_spec_istft", "flowse_op_head_conv",
    "flowse_fp16_overflow", "flowse_rk_lincomb", "flowse_packed_bytes", "flowse_export_packed", "flowse_load_packed",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Context:
    """One libflowse context (weights + workspace) on one CUDA device."""

    def __init__(self, device: Optional[int] = None):
        self._lib = load_library()
        if not torch.cuda.is_available():
            raise FlowseError("flowmse_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        rc = self._lib.flowse_create(C.byref(h), self.device)
        if rc != 0:
            raise FlowseError(self._lib.flowse_last_error(None).decode())
        self._h = h
        self._weights_version = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.flowse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise FlowseError(self._lib.flowse_last_error(self._h).decode() or f"libflowse error {rc}")

    # ---- weights -------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Upload backbone weights given in the reference's state_dict layout (ncsnpp_spec.state_dict_layout)."""
        layout = spec.state_dict_layout()
        descs = (_TensorDesc * len(layout))()
        parts, off = [], 0
        for k, (name, shape) in enumerate(layout):
            if name not in sd:
                raise KeyError(f"state_dict is missing '{name}'")
            t = sd[name].detach().to(device="cpu", dtype=torch.float32).reshape(-1)
            n = 1
            for s in shape:
                n *= s
            if t.numel() != n:
                raise ValueError(f"{name}: {t.numel()} elements, expected {n}")
            descs[k].name = name.encode()
            descs[k].offset = off
            descs[k].numel = n
            parts.append(t)
            off += n
        blob = torch.cat(parts).contiguous()
        self._check(self._lib.flowse_load_weights(self._h, blob.data_ptr(), descs, len(layout)))

    def export_packed(self) -> torch.Tensor:
        """The loaded weights in the library's own packed format (conv weights K-major fp16 hi/lo, small tensors fp32)
        as a uint8 CPU tensor; ``load_packed`` on a fresh context restores them bit for bit."""
        n = int(self._lib.flowse_packed_bytes(self._h))
        if n == 0:
            raise FlowseError("export_packed: no weights loaded")
        blob = torch.empty(n, dtype=torch.uint8)
        self._check(self._lib.flowse_export_packed(self._h, blob.data_ptr(), n))
        return blob

    def load_packed(self, blob: torch.Tensor):
        blob = blob.detach().to(device="cpu", dtype=torch.uint8).contiguous()
        self._check(self._lib.flowse_load_packed(self._h, blob.data_ptr(), blob.numel()))

    # ---- hot path ------------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._check(self._lib.flowse_set_option(self._h, key.encode(), int(value)))

    def rk_lincomb(self, base64, K32, coef, out64=None, out32=None, norm_of=None, rtol=0.0, atol=0.0):
        """v = base64 + sum_s coef[s] * K32[s] (complex128 state, complex64 stages K32 [S_max, ...]); optionally written to
        out64 / out32 and, with norm_of = (ya64, yb64), reduced to sum |v / (atol + rtol max(|ya|, |yb|))|^2 (returned;
        synchronises).  Tensors are CUDA, contiguous; the stage count is len(coef)."""
        S = len(coef)
        ref = base64 if base64 is not None else K32[0]
        n = ref.numel()
        arr = (C.c_double * max(S, 1))(*[float(c) for c in coef])
        acc = C.c_double(0.0)
        ya, yb = norm_of if norm_of is not None else (None, None)
        self._check(self._lib.flowse_rk_lincomb(self._h, _ptr(base64), _ptr(K32), K32[0].numel() if K32 is not None else 0,
                                                arr, S, _ptr(out64), _ptr(out32), _ptr(ya), _ptr(yb), float(rtol),
                                                float(atol), C.byref(acc) if norm_of is not None else None, n, _stream()))
        return acc.value if norm_of is not None else None

    def fp16_overflow(self, reset: bool = True) -> int:
        """Operand values seen outside the fp16 hi/lo range since the last reset (0 = the fp32-parity claim holds).
        Synchronises the device."""
        n = C.c_longlong()
        self._check(self._lib.flowse_fp16_overflow(self._h, C.byref(n), int(reset)))
        return int(n.value)

    def kernel_launches(self) -> int:
        return int(self._lib.flowse_kernel_launches(self._h))

    def workspace_bytes(self, B: int, T: int) -> int:
        n = int(self._lib.flowse_workspace_bytes(self._h, B, T))
        if n == 0:
            self._check(1)
        return n

    @staticmethod
    def _check_spec(t: torch.Tensor, name: str):
        if t.dtype != torch.complex64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous complex64 CUDA tensor")
        if t.dim() != 4 or t.shape[2] != spec.IMAGE_SIZE or t.shape[3] % 64 != 0:
            raise ValueError(f"{name} must have shape [B, C, 256, T] with T % 64 == 0 (pad_spec), got {tuple(t.shape)}")

    def prior_sample(self, y: torch.Tensor, z: torch.Tensor, sigma: float) -> torch.Tensor:
        x = torch.empty_like(y)
        self._check(self._lib.flowse_prior_sample(self._h, y.data_ptr(), z.data_ptr(), float(sigma), x.data_ptr(),
                                                  y.numel(), _stream()))
        return x

    def euler_step(self, x: torch.Tensor, v: torch.Tensor, stepsize: float) -> torch.Tensor:
        out = torch.empty_like(x)
        self._check(self._lib.flowse_euler_step(self._h, x.data_ptr(), v.data_ptr(), float(stepsize), out.data_ptr(),
                                                x.numel(), _stream()))
        return out

    def ncsnpp_forward(self, xy: torch.Tensor, t: torch.Tensor, negate: bool = False) -> torch.Tensor:
        """NCSNpp.forward on a [B,2,256,T] complex64 tensor (channel 0 = state, 1 = condition)."""
        self._check_spec(xy, "x")
        if xy.shape[1] != 2:
            raise ValueError("dnn input must have 2 complex channels (x, y)")
        B, _, F, T = xy.shape
        t = t.to(device=xy.device, dtype=torch.float32).contiguous()
        if t.numel() != B:
            raise ValueError("time_cond must have one entry per batch element")
        out = torch.empty((B, 1, F, T), dtype=torch.complex64, device=xy.device)
        base = xy.data_ptr()
        self._check(self._lib.flowse_ncsnpp_forward(self._h, base, 2 * F * T, base + F * T * 8, 2 * F * T, t.data_ptr(),
                                                    out.data_ptr(), int(negate), B, T, _stream()))
        return out

    def vf_forward(self, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """VFModel.forward(x, t, y) = -dnn(cat([x, y], 1), t) without materialising the concat."""
        self._check_spec(x, "x"); self._check_spec(y, "y")
        B, _, F, T = x.shape
        t = t.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(x)
        self._check(self._lib.flowse_ncsnpp_forward(self._h, x.data_ptr(), F * T, y.data_ptr(), F * T, t.data_ptr(),
                                                    out.data_ptr(), 1, B, T, _stream()))
        return out

    def sample(self, y: torch.Tensor, z: torch.Tensor, timesteps: torch.Tensor, solver: int = 0,
               sigma: float = 0.487, y_prior: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._check_spec(y, "Y"); self._check_spec(z, "z")
        if y_prior is not None:
            self._check_spec(y_prior, "Y_prior")
        B, _, F, T = y.shape
        if isinstance(timesteps, torch.Tensor):              # a device tensor costs a synchronisation here
            timesteps = timesteps.detach().to(device="cpu", dtype=torch.float32).tolist()
        arr = (C.c_float * len(timesteps))(*timesteps)
        out = torch.empty_like(y)
        self._check(self._lib.flowse_sample(self._h, y.data_ptr(), _ptr(y_prior), z.data_ptr(), arr, len(timesteps),
                                            int(solver), float(sigma), out.data_ptr(), B, T, _stream()))
        return out

    # ---- STFT / iSTFT either side of the sampler ----------------------------------
    @staticmethod
    def frames_of(length: int) -> int:
        """Frames torch.stft(center=True, hop 128) yields for `length` samples (data_module.py:163-170)."""
        return 1 + int(length) // 128

    _TRANSFORMS = {"exponent": 0, "log": 1, "none": 2}
    _WINDOWS = {"hann": 0, "sqrthann": 1}

    def _stft_mode(self, transform_type: str, window: str):
        """Select the amplitude transform (data_module.py:149-175) and the STFT window (data_module.py:13-19)."""
        if transform_type not in self._TRANSFORMS:
            raise ValueError(f"transform_type must be one of {sorted(self._TRANSFORMS)}")
        if window not in self._WINDOWS:
            raise NotImplementedError(f"Window type {window} not implemented!")
        mode = (self._TRANSFORMS[transform_type], self._WINDOWS[window])
        if getattr(self, "_stft_mode_set", None) != mode:
            self.set_option("spec_transform", mode[0])
            self.set_option("stft_window", mode[1])
            self._stft_mode_set = mode

    def stft_spec(self, wav: torch.Tensor, lengths, normalize: bool = True, spec_factor: float = 0.15,
                  abs_exponent: float = 0.5, Tpad: Optional[int] = None, transform_type: str = "exponent",
                  window: str = "hann"):
        """wav: fp32 CUDA [B, Lmax] (rows zero-padded to the longest utterance), lengths: samples per utterance.
        Returns (Y complex64 [B,1,256,Tpad], peak fp32 [B]) - the padded model-domain spectrograms evaluate.py:107-115
        builds per file, for the whole ragged batch at once."""
        self._stft_mode(transform_type, window)
        if wav.dtype != torch.float32 or not wav.is_cuda or wav.dim() != 2 or not wav.is_contiguous():
            raise ValueError("wav must be a contiguous fp32 CUDA tensor [B, Lmax]")
        B = wav.shape[0]
        lens = [int(v) for v in lengths]
        if len(lens) != B or max(lens) > wav.shape[1]:
            raise ValueError("lengths must have one entry per row, each <= wav.shape[1]")
        T = self.frames_of(max(lens))
        if Tpad is None:
            Tpad = ((T + 63) // 64) * 64
        Y = torch.empty((B, 1, spec.IMAGE_SIZE, Tpad), dtype=torch.complex64, device=wav.device)
        peak = torch.ones(B, dtype=torch.float32, device=wav.device)
        arr = (C.c_int * B)(*lens)
        self._check(self._lib.flowse_stft_spec(self._h, wav.data_ptr(), wav.stride(0), arr, B, int(normalize),
                                               float(spec_factor), float(abs_exponent), Y.data_ptr(), Tpad,
                                               peak.data_ptr(), _stream()))
        return Y, peak

    def spec_istft(self, X: torch.Tensor, lengths, peak: Optional[torch.Tensor] = None, spec_factor: float = 0.15,
                   abs_exponent: float = 0.5, transform_type: str = "exponent", window: str = "hann") -> torch.Tensor:
        """X: complex64 CUDA [B,1,256,Tpad] -> fp32 [B, max(lengths)] waveforms (VFModel.to_audio per row, times peak)."""
        self._stft_mode(transform_type, window)
        if X.dtype != torch.complex64 or not X.is_cuda or X.dim() != 4 or not X.is_contiguous():
            raise ValueError("X must be a contiguous complex64 CUDA tensor [B,1,256,Tpad]")
        B, Tpad = X.shape[0], X.shape[3]
        lens = [int(v) for v in lengths]
        if len(lens) != B:
            raise ValueError("lengths must have one entry per batch element")
        out = torch.empty((B, max(lens)), dtype=torch.float32, device=X.device)
        arr = (C.c_int * B)(*lens)
        self._check(self._lib.flowse_spec_istft(self._h, X.data_ptr(), Tpad, arr, B, float(spec_factor),
                                                float(abs_exponent), _ptr(peak), out.data_ptr(), out.stride(0),
                                                _stream()))
        return out

    def profile_forward(self):
        """Per-op device times of one network evaluation of the current plan: list of dicts
        {kind, ms, flops, H, W, K, Cout}."""
        import numpy as np
        n_max = 4096
        kinds = np.zeros(n_max, np.int32); ms = np.zeros(n_max, np.float32)
        flops = np.zeros(n_max, np.float64); info = np.zeros(4 * n_max, np.int32)
        n = C.c_int()
        self._check(self._lib.flowse_profile_forward(self._h, n_max, kinds.ctypes.data, ms.ctypes.data,
                                                     flops.ctypes.data, info.ctypes.data, C.byref(n)))
        names = ["misc", "gn_stats", "gn_prep", "conv_gemm", "attention", "small", "temb", "conv_halo"]
        return [dict(kind=names[kinds[k]], ms=float(ms[k]), flops=float(flops[k]), H=int(info[4 * k]),
                     W=int(info[4 * k + 1]), K=int(info[4 * k + 2]), Cout=int(info[4 * k + 3])) for k in range(n.value)]

    # ---- test hooks ----------------------------------------------------------------------------
    def debug_tap(self, module_idx: int, B: int) -> torch.Tensor:
        """Copy of all_modules[module_idx]'s output from the last forward, as NCHW fp32."""
        p, c, h, w = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.flowse_debug_tap(self._h, module_idx, C.byref(p), C.byref(c), C.byref(h), C.byref(w)))
        n = B * h.value * w.value * c.value
        out = torch.empty(n, dtype=torch.float32, device=f"cuda:{self.device}")
        self._check(self._lib.flowse_debug_copy(self._h, p.value, out.data_ptr(), n * 4))
        return out.view(B, h.value, w.value, c.value).permute(0, 3, 1, 2).contiguous()

    def pack_conv_weights(self, w_main: torch.Tensor, w_sc: Optional[torch.Tensor], npad: int):
        """[Cout,Cin,kh,kw] (+[Cout,Cin2,1,1]) fp32 -> (device half buffer [2,Npad,K], wexp)."""
        w_main = w_main.detach().cpu().float().contiguous()
        cout, cin = w_main.shape[0], w_main.shape[1]
        ntaps = w_main.shape[2] * w_main.shape[3]
        cin2 = 0
        if w_sc is not None:
            w_sc = w_sc.detach().cpu().float().contiguous()
            cin2 = w_sc.shape[1]
        K = ntaps * cin + cin2
        out = torch.empty((2, npad, K), dtype=torch.float16, device=f"cuda:{self.device}")
        e = C.c_int()
        rc = self._lib.flowse_pack_conv_weights(w_main.data_ptr(), cout, cin, ntaps, _ptr(w_sc), cin2, npad,
                                                out.data_ptr(), C.byref(e))
        if rc != 0:
            raise FlowseError("pack_conv_weights failed")
        return out, e.value

    def op_gn_prep(self, src1, src2, gamma, beta, mode=0, silu=True, want_x=False, want_f32=False):
        """src: NHWC fp32 [B,H,W,C].  Returns dict with 'A' (fp16 [2,B,Ho,Wo,C]), optional 'X', 'F', 'XF'."""
        B, H, W, C1 = src1.shape
        C2 = 0 if src2 is None else src2.shape[3]
        Ct = C1 + C2
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        dev = src1.device
        A = torch.empty((2, B, Ho, Wo, Ct), dtype=torch.float16, device=dev)
        X = torch.empty_like(A) if want_x else None
        Fo = torch.empty((B, Ho, Wo, Ct), dtype=torch.float32, device=dev) if want_f32 else None
        XF = torch.empty((B, Ho, Wo, Ct), dtype=torch.float32, device=dev) if want_f32 else None
        self._check(self._lib.flowse_op_gn_prep(self._h, src1.data_ptr(), C1, _ptr(src2), C2, gamma.data_ptr(),
                                                beta.data_ptr(), B, H, W, mode, int(silu), A.data_ptr(), _ptr(X),
                                                _ptr(Fo), _ptr(XF), _stream()))
        return dict(A=A, X=X, F=Fo, XF=XF)

    def op_conv_gemm(self, A, Wp, wexp, bias, cout, ntaps=9, X=None, residual=None, div_sqrt2=False, impl=0,
                     bias_bstride=0, out=None):
        """A: fp16 [2,B,H,W,Cin]; Wp: fp16 [2,Npad,K]; returns NHWC fp32 [B,H,W,cout]."""
        _, B, H, W, Cin = A.shape
        Cin2 = 0 if X is None else X.shape[4]
        npad = Wp.shape[1]
        if out is None:
            out = torch.zeros((B, H, W, cout), dtype=torch.float32, device=A.device)
        self._check(self._lib.flowse_op_conv_gemm(self._h, A.data_ptr(), Cin, ntaps, _ptr(X), Cin2, Wp.data_ptr(), npad,
                                                  wexp, bias.data_ptr(), bias_bstride, _ptr(residual), int(div_sqrt2),
                                                  out.data_ptr(), cout, cout, B, H, W, impl, _stream()))
        return out

    def op_head_conv(self, h_nhwc, gamma, beta, weight, bias, prev=None):
        """Pyramid head on NHWC fp32 h; weight in the reference layout [4, C, 3, 3]; prev NHWC [B,H/2,W/2,4] or None."""
        B, H, W, Cc = h_nhwc.shape
        wf = weight.permute(2, 3, 1, 0).reshape(9, Cc, 4).contiguous()          # [tap][C][4]
        out = torch.empty((B, H, W, 4), dtype=torch.float32, device=h_nhwc.device)
        self._check(self._lib.flowse_op_head_conv(self._h, h_nhwc.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                  wf.data_ptr(), bias.data_ptr(), _ptr(prev), out.data_ptr(), B, H, W,
                                                  Cc, _stream()))
        return out

    def op_attention(self, module_idx: int, x_nhwc: torch.Tensor) -> torch.Tensor:
        B, H, W, _ = x_nhwc.shape
        out = torch.empty_like(x_nhwc)
        self._check(self._lib.flowse_op_attention(self._h, module_idx, x_nhwc.data_ptr(), out.data_ptr(), B, H, W,
                                                  _stream()))
        return out
