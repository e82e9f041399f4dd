"""BASELINE.json configs as GPU parity cases: sizes the CPU oracle finishes in seconds are compared with it directly;
the full-size cases use size-independent properties (batch consistency, chained-step consistency)."""
import numpy as np
import pytest
import torch

from oracle import ncsnpp_oracle as orc

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


@pytest.fixture(scope="module")
def ctx(synthetic_sd):
    from flowmse_b200.lib import Context
    c = Context(0)
    c.load_state_dict(synthetic_sd)
    yield c
    c.close()


def _rand_c(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(scale * torch.randn(*shape, 2, generator=g))


def _frac_bad(a, b):
    a, b = torch.view_as_real(a.cpu()), torch.view_as_real(b.cpu())
    bad = (a - b).abs() > ATOL + RTOL * b.abs()
    return bad.float().mean().item(), (a - b).abs().max().item()


def test_config1_wav_to_wav_N1(synthetic_sd):
    """configs[0]: 4 s synthetic wav, N=1 Euler through the reference-facing API (VFModel.enhance), vs the oracle driven
    with the same prior draw."""
    from flowmse_b200.model import VFModel
    from flowmse_b200.util.other import pad_spec
    model = VFModel(backbone="ncsnpp", ode="flowmatching")
    model.dnn.load_state_dict(synthetic_sd, strict=True)
    model.eval()
    torch.manual_seed(0)
    n = 64000
    tg = torch.arange(n) / 16000.0
    wav = 0.1 * torch.randn(1, n) + 0.5 * torch.sin(2 * np.pi * 220 * tg) + 0.25 * torch.sin(2 * np.pi * 1320 * tg)
    wav = (wav / wav.abs().max()).cuda()
    torch.manual_seed(77)
    out = model.enhance(wav, N=1)
    assert out.shape[-1] == n and torch.isfinite(out).all()
    # same pipeline with the oracle in the middle
    Y = pad_spec(torch.unsqueeze(model._forward_transform(model._stft(wav / wav.abs().max())), 0)).contiguous()
    assert Y.shape[-1] == 512
    torch.manual_seed(77)
    z = torch.randn_like(Y)
    x_ref = orc.sample(synthetic_sd, Y.cpu(), z.cpu(), 1)
    torch.manual_seed(77)
    x = model.enhance_spec(Y, N=1)                       # the spectrogram the sampler hands to to_audio
    frac, mx = _frac_bad(x, x_ref)
    assert frac == 0.0, (frac, mx)                       # north-star tolerance holds in the spectrogram domain
    # waveform: the magnitude expansion |X|^2 and the iSTFT mix bins, so compare in relative L2
    ref = model.to_audio(x_ref.cuda().squeeze(), n) * wav.abs().max()
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1e-4, rel


def test_ragged_T640_vs_oracle(ctx, synthetic_sd):
    """T = 640 (not a power of two: widths 640 ... 10 down the U-Net exercise partial tiles at the low resolutions)."""
    Y, z = _rand_c((1, 1, 256, 640), 21, 0.3), _rand_c((1, 1, 256, 640), 22, np.sqrt(0.5))
    x_ref = orc.sample(synthetic_sd, Y, z, 1)
    x = ctx.sample(Y.cuda(), z.cuda(), torch.linspace(1.0, 0.03, 1), solver=0, sigma=0.487)
    frac, mx = _frac_bad(x, x_ref)
    assert frac == 0.0, (frac, mx)


def test_config3_batch16_heun25_batch_consistency(ctx):
    """configs[2]: B=16, T=512, N=25 Heun (49 chained NFEs).  The CPU oracle would need hours, so the full-size case is
    checked through batch consistency: element i of the batched run vs the same utterance sampled alone (another plan:
    other tile / split-K choices, i.e. another valid fp32 evaluation order), and every output finite.

    Tolerance: after ONE NFE the two orders agree to 4e-5 max abs (rms 8e-6); through 49 NFEs the sampler's own dynamics
    amplify that noise for sensitive utterances (measured with tools/diag_batch.py on B200: element 0 -> 0.35 % of the
    bins outside rtol 1e-3 / atol 1e-4, max abs 2.6e-3; element 11 -> 0.003 %, 3.0e-4; the per-tap kernel with three
    rotating accumulators shows the same growth from a 2.5x lower start).  Any two fp32 implementations - the reference
    on two BLAS back ends included - separate like this, so the element-wise north-star tolerance is asserted for
    >= 99 % of the bins with a bound on the worst bin; the short chains (configs 1-2) are held to 100 %."""
    B, T, N = 16, 512, 25
    Y, z = _rand_c((B, 1, 256, T), 31, 0.3).cuda(), _rand_c((B, 1, 256, T), 32, np.sqrt(0.5)).cuda()
    ts = torch.linspace(1.0, 0.03, N)
    xb = ctx.sample(Y, z, ts, solver=1, sigma=0.487)
    torch.cuda.synchronize()
    assert torch.isfinite(torch.view_as_real(xb)).all()
    for i in (0, 11):
        xi = ctx.sample(Y[i:i + 1].contiguous(), z[i:i + 1].contiguous(), ts, solver=1, sigma=0.487)
        frac, mx = _frac_bad(xb[i:i + 1], xi)
        assert frac <= 1e-2 and mx < 1e-2, (i, frac, mx)
    # one NFE (N=1 Euler) at the same size: 100 % inside the tolerance
    x1b = ctx.sample(Y, z, torch.linspace(1.0, 0.03, 1), solver=0, sigma=0.487)
    x1 = ctx.sample(Y[3:4].contiguous(), z[3:4].contiguous(), torch.linspace(1.0, 0.03, 1), solver=0, sigma=0.487)
    frac, mx = _frac_bad(x1b[3:4], x1)
    assert frac == 0.0 and mx < 1e-4, (frac, mx)


def test_config2_N5_euler_chained_steps_property(ctx):
    """configs[1] at full size (B=1, T=512, N=5): flowse_sample (one call, graph replay, fused final kernel) equals the
    reference's own loop structure driven step by step through the C ABI (vf_forward + euler_step per step)."""
    Y, z = _rand_c((1, 1, 256, 512), 41, 0.3).cuda(), _rand_c((1, 1, 256, 512), 42, np.sqrt(0.5)).cuda()
    ts = torch.linspace(1.0, 0.03, 5)
    x_fused = ctx.sample(Y, z, ts, solver=0, sigma=0.487)
    x = ctx.prior_sample(Y, z, 0.487)
    for i in range(5):
        step = ts[i] - ts[i + 1] if i != 4 else ts[-1]
        v = ctx.vf_forward(x, torch.full((1,), float(ts[i]), device="cuda"), Y)
        x = ctx.euler_step(x, v, float(step))
    frac, mx = _frac_bad(x_fused, x)
    assert frac == 0.0, (frac, mx)


def test_longest_bucket_T1280_batch_consistency(ctx):
    """The longest utterances of the VoiceBank-DEMAND-shaped set (10 s -> T = 1280 frames, 256 x 1280 pixels at the top
    level, 4 x 20 at the bottom: partial 128-pixel tiles on every low-resolution level) and the shortest (T = 64): one NFE
    (N = 1 Euler) of a batch of two vs each utterance alone - another plan, other tile / split-K cluster choices - must
    agree within the north-star tolerance on 100 % of the bins, and everything must be finite."""
    for T in (1280, 64):
        Y, z = _rand_c((2, 1, 256, T), 51, 0.3).cuda(), _rand_c((2, 1, 256, T), 52, np.sqrt(0.5)).cuda()
        ts = torch.linspace(1.0, 0.03, 1)
        xb = ctx.sample(Y, z, ts, solver=0, sigma=0.487)
        assert torch.isfinite(torch.view_as_real(xb)).all()
        for i in (0, 1):
            xi = ctx.sample(Y[i:i + 1].contiguous(), z[i:i + 1].contiguous(), ts, solver=0, sigma=0.487)
            frac, mx = _frac_bad(xb[i:i + 1], xi)
            assert frac == 0.0 and mx < 2e-4, (T, i, frac, mx)
