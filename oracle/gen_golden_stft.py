"""Generate tests/golden/stft_roundtrip.npz with the REFERENCE's own STFT helpers (imported from /root/reference):
SpecsDataModule.stft / spec_fwd / spec_back / istft (flowmse/data_module.py:149-205) exactly as VFModel._stft,
_forward_transform, _backward_transform, to_audio delegate to them (flowmse/model.py:190-203), plus pad_spec
(flowmse/util/other.py:83-90) and evaluate.py's peak normalisation (evaluate.py:109-110).

Run here (CPU container):  python oracle/gen_golden_stft.py
Test infrastructure only: the vectors pin flowse_stft_spec / flowse_spec_istft.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from gen_golden import install_stubs, REF   # noqa: E402


def main():
    install_stubs()
    sys.path.insert(0, REF)
    torch.set_grad_enabled(False)
    from flowmse.data_module import SpecsDataModule
    from flowmse.util.other import pad_spec

    dm = SpecsDataModule(base_dir="")           # defaults: n_fft 510, hop 128, hann, exponent 0.5, factor 0.15
    out = {}
    for name, n, seed in (("a", 9000, 0), ("b", 5003, 1)):      # 71 and 40 frames; "b" has a ragged last hop
        g = torch.Generator().manual_seed(seed)
        tg = torch.arange(n) / 16000.0
        wav = 0.2 * torch.randn(1, n, generator=g) + 0.5 * torch.sin(2 * np.pi * 330 * tg) + 0.2 * torch.sin(2 * np.pi * 2500 * tg)
        norm = wav.abs().max()
        Y = torch.unsqueeze(dm.spec_fwd(dm.stft(wav / norm)), 0)         # [1,1,256,T]
        Yp = pad_spec(Y)
        # an "enhanced" spectrogram to invert: a smooth, deterministic modification of Y (keeps magnitudes realistic)
        X = Yp * (0.8 + 0.2j)
        x_hat = dm.istft(dm.spec_back(X.squeeze()), n) * norm
        out[f"wav_{name}"] = wav.numpy()
        out[f"norm_{name}"] = np.array([norm.item()], dtype=np.float32)
        out[f"Y_{name}"] = torch.view_as_real(Yp).numpy()
        out[f"X_{name}"] = torch.view_as_real(X).numpy()
        out[f"xhat_{name}"] = x_hat.numpy()
        print(name, "wav", tuple(wav.shape), "Y", tuple(Y.shape), "->", tuple(Yp.shape), "x_hat", tuple(x_hat.shape))
    if "--variants-only" not in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stft_roundtrip.npz"), **out)

    # the non-default transform_type / window choices of SpecsDataModule (data_module.py:13-19, 149-175), one short
    # utterance each -> tests/golden/stft_variants.npz
    var = {}
    n = 5003
    g = torch.Generator().manual_seed(2)
    tg = torch.arange(n) / 16000.0
    wav = 0.2 * torch.randn(1, n, generator=g) + 0.5 * torch.sin(2 * np.pi * 440 * tg)
    norm = wav.abs().max()
    var["wav"] = wav.numpy()
    var["norm"] = np.array([norm.item()], dtype=np.float32)
    for transform_type, window in (("log", "hann"), ("exponent", "sqrthann"), ("none", "sqrthann"), ("log", "sqrthann")):
        dmv = SpecsDataModule(base_dir="", transform_type=transform_type, window=window)
        Yp = pad_spec(torch.unsqueeze(dmv.spec_fwd(dmv.stft(wav / norm)), 0))
        X = Yp * (0.8 + 0.2j)
        x_hat = dmv.istft(dmv.spec_back(X.squeeze()), n) * norm
        key = f"{transform_type}_{window}"
        var[f"Y_{key}"] = torch.view_as_real(Yp).numpy()
        var[f"xhat_{key}"] = x_hat.numpy()
        print(key, "Y", tuple(Yp.shape), "max|Y|", Yp.abs().max().item(), "max|x_hat|", x_hat.abs().max().item())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stft_variants.npz"), **var)


if __name__ == "__main__":
    main()
