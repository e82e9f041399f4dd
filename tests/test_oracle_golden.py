"""Pin the CPU oracle against vectors produced by the reference itself (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ncsnpp_oracle as orc


def _c(a):
    return torch.view_as_complex(torch.from_numpy(np.ascontiguousarray(a)))


def test_schedule_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    for N in (1, 2, 5, 25, 30):
        ts, steps = orc.schedule(N)
        assert ts.numpy().tobytes() == g[f"t{N}"].tobytes()
        assert steps.numpy().tobytes() == g[f"s{N}"].tobytes()
    # fp32 hex values of the N=5 schedule
    want = ["0x1.000000p+0", "0x1.83d70ap-1", "0x1.07ae14p-1", "0x1.170a3ep-2", "0x1.eb851ep-6"]
    got = orc.schedule(5)[0].numpy()
    assert [float.fromhex(w) for w in want] == [float(v) for v in got]
    assert float(orc.schedule(1)[1][0]) == 1.0


def test_fir_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "fir.npz"))
    x = torch.from_numpy(g["x"])
    assert torch.allclose(orc.fir_upsample2(x), torch.from_numpy(g["up"]), atol=1e-6, rtol=0)
    assert torch.allclose(orc.fir_downsample2(x), torch.from_numpy(g["down"]), atol=1e-6, rtol=0)


def test_forward_matches_reference(golden_dir, synthetic_sd):
    g = np.load(os.path.join(golden_dir, "forward_T64.npz"))
    x, t, v_ref = _c(g["x"]), torch.from_numpy(g["t"]), _c(g["v"])
    with torch.no_grad():
        v = orc.ncsnpp_forward(synthetic_sd, x, t)
    # north_star tolerance: rtol 1e-3 / atol 1e-4 fp32
    assert torch.allclose(torch.view_as_real(v), torch.view_as_real(v_ref), rtol=1e-3, atol=1e-4)
    assert (v - v_ref).abs().max().item() < 2e-3  # measured 2.5e-4 on |v| up to 220


def test_sampler_matches_reference(golden_dir, synthetic_sd):
    g = np.load(os.path.join(golden_dir, "sampler_T64.npz"))
    Y, z = _c(g["Y"]), _c(g["z"])
    x1 = orc.sample(synthetic_sd, Y, z, 1)
    assert torch.allclose(torch.view_as_real(x1), torch.from_numpy(g["x_euler_N1"]), rtol=1e-3, atol=1e-4)


def test_pad_spec():
    y = torch.zeros(1, 1, 256, 501, dtype=torch.complex64)
    assert orc.pad_spec(y).shape[-1] == 512
    assert orc.pad_spec(orc.pad_spec(y)).shape[-1] == 512


@pytest.mark.parametrize("name", ["a", "b"])
def test_oracle_stft_istft_match_reference_golden(golden_dir, name):
    """oracle.stft_spec / spec_istft vs the vectors the reference's SpecsDataModule produced (gen_golden_stft.py)."""
    g = np.load(os.path.join(golden_dir, "stft_roundtrip.npz"))
    wav = torch.from_numpy(g[f"wav_{name}"])
    Y, peak = orc.stft_spec(wav)
    assert torch.equal(torch.view_as_real(Y), torch.from_numpy(g[f"Y_{name}"]))
    assert float(peak) == float(g[f"norm_{name}"][0])
    X = torch.view_as_complex(torch.from_numpy(g[f"X_{name}"]).contiguous())
    xh = orc.spec_istft(X, wav.shape[1], peak)
    assert torch.allclose(xh, torch.from_numpy(g[f"xhat_{name}"]), rtol=0, atol=1e-7)
