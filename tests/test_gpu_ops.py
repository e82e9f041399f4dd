"""Op-level GPU parity tests: every kernel family called through the C ABI (ctypes) and compared with a plain
fp32/fp64 torch restatement of the same op on the same seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ncsnpp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from flowmse_b200.lib import Context
    c = Context(0)
    yield c
    c.close()


def _split(a):
    hi = a.half()
    lo = (a - hi.float()).half()
    return torch.stack([hi, lo]).contiguous()


def _join(s):
    return s[0].float() + s[1].float()


def test_prior_and_euler_bit_exact(ctx):
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.view_as_complex(torch.randn(2, 1, 256, 64, 2, device="cuda", generator=g))
    z = torch.view_as_complex(torch.randn(2, 1, 256, 64, 2, device="cuda", generator=g))
    x = ctx.prior_sample(y, z, 0.487)
    std = torch.full((2,), 0.487, device="cuda")
    ref = y + z * std[:, None, None, None]
    assert torch.equal(torch.view_as_real(x), torch.view_as_real(ref))
    v = torch.view_as_complex(torch.randn(2, 1, 256, 64, 2, device="cuda", generator=g))
    step = torch.tensor(0.2425, device="cuda")
    ref2 = x + v * (-step)
    out = ctx.euler_step(x, v, float(step))
    assert torch.equal(torch.view_as_real(out), torch.view_as_real(ref2))
    # odd element count exercises the scalar tail
    y1 = y.reshape(-1)[:1001].contiguous(); z1 = z.reshape(-1)[:1001].contiguous()
    x1 = ctx.prior_sample(y1, z1, 0.487)
    assert torch.equal(torch.view_as_real(x1), torch.view_as_real(y1 + z1 * torch.tensor(0.487, device="cuda")))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("C1,C2", [(128, 0), (256, 128), (256, 256)])
def test_gn_prep_matches_torch(ctx, mode, C1, C2):
    if mode != 0 and C2:
        pytest.skip("resampling blocks never take a concatenated input")
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W = 2, 8, 12
    s1 = torch.randn(B, H, W, C1, device="cuda", generator=g) * 1.7 + 0.3
    s2 = torch.randn(B, H, W, C2, device="cuda", generator=g) if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    r = ctx.op_gn_prep(s1, s2, gamma, beta, mode=mode, silu=True, want_x=True, want_f32=True)
    x = s1 if s2 is None else torch.cat([s1, s2], dim=3)
    x = x.permute(0, 3, 1, 2).double().cpu()
    h = F.group_norm(x, 32, gamma.double().cpu(), beta.double().cpu(), eps=1e-6)
    h = h * torch.sigmoid(h)
    if mode == 1:
        h, x = orc.fir_downsample2(h), orc.fir_downsample2(x)
    elif mode == 2:
        h, x = orc.fir_upsample2(h), orc.fir_upsample2(x)
    h = h.permute(0, 2, 3, 1).float(); x = x.permute(0, 2, 3, 1).float()
    assert torch.allclose(r["F"].cpu(), h, rtol=1e-5, atol=2e-6)
    assert torch.allclose(r["XF"].cpu(), x, rtol=1e-5, atol=2e-6)
    # hi + lo reproduces the fp32 value to ~2^-22 relative
    assert torch.allclose(_join(r["A"]).cpu(), r["F"].cpu(), rtol=5e-7, atol=1e-7)
    assert torch.allclose(_join(r["X"]).cpu(), r["XF"].cpu(), rtol=5e-7, atol=1e-7)


CONV_CASES = [
    # B, H, W, Cin, Cout, ntaps, Cin2, residual, div
    (1, 8, 16, 128, 128, 9, 0, True, True),
    (2, 4, 2, 256, 256, 9, 512, False, True),
    (1, 16, 40, 128, 256, 9, 0, False, False),
    (1, 32, 64, 384, 128, 9, 384, False, True),
    (2, 16, 8, 256, 4, 9, 0, False, False),
    (1, 8, 8, 128, 128, 1, 0, False, False),
]


def _conv_case(ctx, case, impl):
    B, H, W, Cin, Cout, ntaps, Cin2, use_res, div = case
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    A = _split(a)
    ks = 3 if ntaps == 9 else 1
    w = torch.randn(Cout, Cin, ks, ks, generator=torch.Generator().manual_seed(1)) / np.sqrt(Cin * ntaps)
    wsc = xs = X = None
    if Cin2:
        wsc = torch.randn(Cout, Cin2, 1, 1, generator=torch.Generator().manual_seed(2)) / np.sqrt(Cin2)
        xs = torch.randn(B, H, W, Cin2, device="cuda", generator=g) * 3.0
        X = _split(xs)
    npad = 16 if Cout < 16 else ((Cout + 127) // 128) * 128
    Wp, wexp = ctx.pack_conv_weights(w, wsc, npad)
    bias = torch.randn(B, Cout, device="cuda", generator=g)
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g) if use_res else None
    out = ctx.op_conv_gemm(A, Wp, wexp, bias, Cout, ntaps=ntaps, X=X, residual=res, div_sqrt2=div, impl=impl,
                           bias_bstride=Cout)
    torch.cuda.synchronize()
    # fp64 reference on exactly the operands the kernel sees (hi+lo activations, fp32 weights)
    ref = F.conv2d(_join(A).permute(0, 3, 1, 2).double().cpu(), w.double(), padding=ks // 2)
    if Cin2:
        ref = ref + F.conv2d(_join(X).permute(0, 3, 1, 2).double().cpu(), wsc.double())
    ref = ref + bias.double().cpu()[:, :, None, None]
    if use_res:
        ref = ref + res.permute(0, 3, 1, 2).double().cpu()
    if div:
        ref = ref / np.sqrt(2.0)
    ref = ref.permute(0, 2, 3, 1).float()
    err = (out.cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_gemm_simt(ctx, case):
    err, scale = _conv_case(ctx, case, impl=1)
    assert err < 2e-5 * max(1.0, scale), (err, scale)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_gemm_tcgen05(ctx, case):
    err, scale = _conv_case(ctx, case, impl=0)
    assert err < 2e-5 * max(1.0, scale), (err, scale)


HALO_CASES = [
    # B, H, W, Cin, Cout, ntaps, Cin2, residual, div       (halo kernel: 3x3, H % 16 == 0, W % 8 == 0)
    (1, 16, 8, 128, 128, 9, 0, True, True),                # one tile, image == tile (all four borders padded)
    (1, 16, 40, 128, 256, 9, 0, False, False),             # two N tiles
    (1, 32, 64, 384, 128, 9, 384, False, True),            # folded 1x1 shortcut chunks
    (2, 16, 8, 256, 4, 9, 0, False, False),                # pyramid head (BN = 16), batch 2
    (1, 128, 160, 64, 128, 9, 0, True, True),              # 160 tiles > 148 CTAs: persistent loop, TMEM double buffer
    (2, 64, 96, 128, 256, 9, 128, False, True),            # 192 tiles, shortcut, batch 2
]


@pytest.mark.parametrize("impl", [2, 3, 4])
@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo(ctx, case, impl):
    """impl 2: halo kernel with one main accumulator (double-buffered TMEM); impl 3: three rotating main accumulators;
    impl 4: CTA pairs (tcgen05.mma.cta_group::2, each CTA stages half of the weight block)."""
    err, scale = _conv_case(ctx, case, impl=impl)
    assert err < 2e-5 * max(1.0, scale), (err, scale)


@pytest.mark.parametrize("B,H,W,mod", [(1, 16, 32, 21), (2, 16, 8, 23), (2, 4, 2, 33), (1, 4, 8, 33), (1, 16, 20, 50),
                                       (4, 16, 32, 21)])
def test_attention_block_matches_oracle(synthetic_sd, B, H, W, mod):
    """AttnBlockpp (layerspp.py:62-91) through flowse_op_attention vs the CPU oracle; token counts 512 / 128 / 8 / 32 /
    320 cover the ragged (non multiple-of-8) row blocks and, with the batch sizes, every rows-per-CTA variant of the two
    kernels (core: 4 rows at B=1 x 512 tokens, 2 rows for the small blocks, 8 rows at B=4 x 512; QKV: 8 and 2 rows)."""
    from flowmse_b200.lib import Context
    c = Context(0)
    c.load_state_dict(synthetic_sd)
    try:
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, 256, H, W, generator=g) * 1.3
        with torch.no_grad():
            ref = orc.attnblock(synthetic_sd, f"all_modules.{mod}.", x)
        out = c.op_attention(mod, x.permute(0, 2, 3, 1).contiguous().cuda()).permute(0, 3, 1, 2).cpu()
        assert torch.allclose(out, ref, rtol=1e-4, atol=2e-5), (out - ref).abs().max().item()
    finally:
        c.close()
