"""Host-side logic that must hold without a GPU: the C-ABI library loads and exports every symbol include/flowse.h
declares, the reference-facing plugin surface (registries, schedule, checkpoint layout), the loud failure without a
CUDA device, and the utterance sharding over a world_size-2 gloo process group."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "flowse.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flowse_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from flowmse_b200 import lib
    handle = lib.load_library()
    declared = _header_symbols()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(handle, sym), f"libflowse.so does not export {sym} (declared in include/flowse.h)"
    # the ctypes table in lib.py covers the whole header
    assert sorted(lib.EXPORTED_SYMBOLS) == declared


def test_every_option_key_is_documented_in_the_header():
    """Every key flowse_set_option accepts (engine.cu) is described in include/flowse.h."""
    import re
    eng = open(os.path.join(ROOT, "flowmse_b200", "csrc", "engine.cu")).read()
    body = eng[eng.index("int flowse_set_option("):]
    body = body[:body.index("\n}\n")]
    keys = set(re.findall(r'k == "([a-z_0-9]+)"', body))
    assert {"conv_impl", "fuse_prep", "graph", "pdl", "whole_graph", "fork", "stft_window", "spec_transform"} <= keys
    hdr = open(os.path.join(ROOT, "include", "flowse.h")).read()
    missing = [k for k in keys if f'"{k}"' not in hdr]
    assert not missing, missing


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (create returns an error, Context raises)."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flowmse_b200 import lib
    handle = lib.load_library()
    h = ctypes.c_void_p()
    assert handle.flowse_create(ctypes.byref(h), 0) != 0
    assert b"no CPU fallback" in handle.flowse_last_error(None)
    with pytest.raises(lib.FlowseError):
        lib.Context(0)
    from flowmse_b200.runtime import get_context
    with pytest.raises(lib.FlowseError):
        get_context("cpu")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flowmse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


# ---------------------------------------------------------------------------------------------------------------
# plugin surface
# ---------------------------------------------------------------------------------------------------------------
def test_registries_match_reference_contract():
    from flowmse_b200.sampling import ODEsolverRegistry
    from flowmse_b200.odes import ODERegistry
    from flowmse_b200.backbones import BackboneRegistry
    assert "euler" in ODEsolverRegistry.get_all_names()
    assert "flowmatching" in ODERegistry.get_all_names()
    assert "ncsnpp" in BackboneRegistry.get_all_names()
    with pytest.raises(ValueError):                       # registry.py:27-30
        ODEsolverRegistry.get_by_name("rk45")
    from flowmse_b200.sampling import get_white_box_solver
    with pytest.raises(ValueError):
        get_white_box_solver("nope", None, None, Y=torch.zeros(1, 1, 256, 64, dtype=torch.complex64))


def test_schedule_matches_golden_bits(golden_dir):
    from flowmse_b200.sampling import timesteps_and_stepsizes
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    for N in (1, 2, 5, 25, 30):
        ts, steps = timesteps_and_stepsizes(N)
        assert ts.numpy().tobytes() == g[f"t{N}"].tobytes()
        assert steps.numpy().tobytes() == g[f"s{N}"].tobytes()


def test_prior_std_is_fp32_sigma_max():
    from flowmse_b200.odes import FLOWMATCHING
    assert FLOWMATCHING().prior_std() == float(np.float32(0.487))
    assert FLOWMATCHING(sigma_min=0.1, sigma_max=0.5).prior_std() == float(np.float32(0.0) * np.float32(0.1) + np.float32(0.5))


def test_checkpoint_layout_roundtrip(synthetic_sd):
    from flowmse_b200 import checkpoint as ck, ncsnpp_spec as spec
    layout = spec.state_dict_layout()
    assert len(layout) == 647
    want = [l.split()[0] for l in open(os.path.join(ROOT, "tests", "golden", "state_dict_layout.txt")) if l.strip()]
    assert [n for n, _ in layout] == want                  # names in the reference's own state_dict order
    assert spec.num_params() == 65_590_822
    ema = {k: v + 1.0 for k, v in synthetic_sd.items()}
    for frozen in (True, False):                           # torch_ema may or may not track the frozen Fourier W
        c = ck.make_lightning_checkpoint(synthetic_sd, ema, include_frozen_in_ema=frozen)
        live = ck.backbone_state_from_checkpoint(c, use_ema=False)
        shadow = ck.backbone_state_from_checkpoint(c, use_ema=True)
        k = "all_modules.4.Conv_0.weight"
        assert torch.equal(live[k], synthetic_sd[k]) and torch.equal(shadow[k], ema[k])
        assert torch.equal(shadow["all_modules.0.W"], ema["all_modules.0.W"] if frozen else synthetic_sd["all_modules.0.W"])
    blob = ck.flatten_state_dict(synthetic_sd)
    back = ck.unflatten_state_dict(blob)
    assert all(torch.equal(back[n], synthetic_sd[n]) for n, _ in layout)
    bad = ck.make_lightning_checkpoint(synthetic_sd)
    bad["ema"]["shadow_params"] = bad["ema"]["shadow_params"][:10]
    with pytest.raises(ValueError):
        ck.backbone_state_from_checkpoint(bad)


def test_pad_spec_contract():
    from flowmse_b200.util.other import pad_spec
    y = torch.zeros(1, 1, 256, 501, dtype=torch.complex64)
    assert pad_spec(y).shape[-1] == 512 and pad_spec(pad_spec(y)).shape[-1] == 512
    assert pad_spec(torch.zeros(1, 1, 256, 126, dtype=torch.complex64)).shape[-1] == 128


# ---------------------------------------------------------------------------------------------------------------
# sharding
# ---------------------------------------------------------------------------------------------------------------
def test_lpt_assignment_and_buckets():
    from flowmse_b200.sharding import lpt_assign, bucket_by_length
    lengths = [512, 128, 1280, 256, 256, 640, 128, 64, 1024, 512]
    parts = lpt_assign(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) <= 1280                              # the longest utterance bounds the optimum here
    assert lpt_assign(lengths, 4) == parts                 # deterministic
    assert lpt_assign([], 3) == [[], [], []]
    b = bucket_by_length([0, 3, 4, 9], lengths, max_batch=1)
    assert b == [[3], [4], [0], [9]]
    b = bucket_by_length([0, 3, 4, 9], lengths, max_batch=8)
    assert b == [[3, 4], [0, 9]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_worker(rank, world, port, ret):
    import torch.distributed as dist
    from flowmse_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # one weight broadcast: rank 0 owns the blob
        blob = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
        sharding.broadcast_weights(blob)
        assert torch.equal(blob, torch.arange(1000, dtype=torch.float32))
        # ragged utterances, identical list on every rank; the "sampler" is a stand-in that tags each bin
        g = torch.Generator().manual_seed(3)
        lengths = [128, 64, 192, 64, 128, 256, 64]
        specs = [torch.view_as_complex(torch.randn(1, 8, T, 2, generator=g)) for T in lengths]
        calls = []

        def fake_enhance(Y):
            calls.append(tuple(Y.shape))
            return Y * (2.0 + 0j) + 1.0

        out = sharding.enhance_sharded(specs, fake_enhance, torch.device("cpu"), max_batch=2)
        assert len(out) == len(specs)
        for o, s in zip(out, specs):
            assert o.shape == s.shape and torch.equal(o, s * (2.0 + 0j) + 1.0)
        mine = sharding.lpt_assign(lengths, world)[rank]
        assert sum(c[0] for c in calls) == len(mine)        # this rank only ran its own share
        ret[rank] = sum(lengths[i] for i in mine)
    finally:
        dist.destroy_process_group()


def test_enhance_sharded_gloo_world2():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_sharded_worker, args=(world, port, ret), nprocs=world, join=True)
        loads = dict(ret)
    assert sorted(loads) == [0, 1]
    assert abs(loads[0] - loads[1]) <= 128                 # LPT balance on [128,64,192,64,128,256,64]


def test_config1_stft_plumbing_matches_reference_golden(golden_dir):
    """BASELINE.json configs[0]: 4 s synthetic wav -> STFT (n_fft 510, hop 128, hann) -> |X|^0.5 e^{j arg X} * 0.15 ->
    pad_spec, against values produced by the reference's own _stft/_forward_transform/pad_spec (oracle/gen_golden.py)."""
    from flowmse_b200.model import SpecTransform
    from flowmse_b200.util.other import pad_spec
    g = np.load(os.path.join(golden_dir, "stft_cfg1.npz"))
    torch.manual_seed(0)
    n = 64000
    tgrid = torch.arange(n) / 16000.0
    wav = 0.1 * torch.randn(1, n) + 0.5 * torch.sin(2 * np.pi * 220 * tgrid) + 0.25 * torch.sin(2 * np.pi * 1320 * tgrid)
    wav = wav / wav.abs().max()
    assert np.array_equal(wav.numpy()[:, :4096], g["wav"])           # same synthetic input as the generator script
    st = SpecTransform()
    Y = torch.unsqueeze(st.spec_fwd(st.stft(wav)), 0)
    assert list(Y.shape) == list(g["shape"]) == [1, 1, 256, 501]
    Yp = pad_spec(Y)
    assert list(Yp.shape) == list(g["padded_shape"]) == [1, 1, 256, 512]
    assert torch.allclose(torch.view_as_real(Yp[0, 0, :, :8]), torch.from_numpy(g["Y_head"]), rtol=1e-5, atol=1e-6)
    # round trip of the transform pair and of the STFT pair (to_audio path, model.py:190-203)
    back = st.istft(st.spec_back(st.spec_fwd(st.stft(wav))), length=n)
    assert torch.allclose(back, wav, atol=2e-4)


@pytest.mark.parametrize("transform_type,window", [("log", "hann"), ("exponent", "sqrthann"), ("none", "sqrthann")])
def test_spec_transform_variants_match_reference_golden(golden_dir, transform_type, window):
    """The torch mirror of SpecsDataModule's other transform_type / window settings (data_module.py:13-19, 149-175)
    against vectors from the reference (oracle/gen_golden_stft.py -> stft_variants.npz)."""
    from flowmse_b200.model import SpecTransform
    from flowmse_b200.util.other import pad_spec
    g = np.load(os.path.join(golden_dir, "stft_variants.npz"))
    key = f"{transform_type}_{window}"
    wav = torch.from_numpy(g["wav"])
    norm = torch.from_numpy(g["norm"])
    st = SpecTransform(transform_type=transform_type, window=window)
    Yp = pad_spec(torch.unsqueeze(st.spec_fwd(st.stft(wav / norm)), 0))
    Yg = torch.view_as_complex(torch.from_numpy(g[f"Y_{key}"]).contiguous())
    assert torch.allclose(torch.view_as_real(Yp), torch.view_as_real(Yg), rtol=1e-5, atol=1e-6)
    x_hat = st.istft(st.spec_back((Yg * (0.8 + 0.2j)).squeeze()), wav.shape[1]) * norm
    assert torch.allclose(x_hat, torch.from_numpy(g[f"xhat_{key}"]), rtol=1e-5, atol=1e-6)


# ---- evaluate driver: host logic ---------------------------------------------------------------
def test_evaluate_cli_matches_reference_arguments():
    """Same argument names / defaults as /root/reference/evaluate.py:27-43."""
    from flowmse_b200.evaluate import build_parser
    a = build_parser().parse_args(["--folder_destination", "/tmp/x", "--test_dir", "/data"])
    assert (a.odesolver_type, a.odesolver, a.reverse_starting_point, a.last_eval_point, a.N, a.N_mid) == \
           ("white", "euler", 1.0, 0.03, 5, 0)
    assert a.ckpt is None and a.test_dir == "/data"


def test_evaluate_batches_and_metrics(tmp_path):
    from flowmse_b200 import evaluate as ev
    # frames / padded frames as torch.stft(center=True, hop 128) + pad_spec give them
    assert ev.frames_of(64000) == 501 and ev.padded_frames(64000) == 512
    assert ev.padded_frames(128 * 63) == 64 and ev.padded_frames(128 * 64) == 128
    n = [64000, 64000, 9000, 30000, 64000, 9100]
    batches = ev.make_batches(range(len(n)), n, max_batch_frames=1024)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(len(n)))
    for b in batches:
        assert len({ev.padded_frames(n[i]) for i in b}) == 1
        assert len(b) == 1 or sum(ev.padded_frames(n[i]) for i in b) <= 1024
    assert batches[0] == [0, 1]                      # longest bucket first, two 512-frame utterances per batch
    # SI-SDR family (utils.py:10-36) on a hand-computable case: orthogonal target / noise / artefact directions
    s = np.zeros(8); s[0] = 1.0
    noise = np.zeros(8); noise[1] = 1.0
    art = np.zeros(8); art[2] = 0.1
    sdr, sir, sar = ev.energy_ratios(2.0 * s + 0.5 * noise + art, s, noise)
    assert abs(sdr - 10 * np.log10(4 / 0.26)) < 1e-9 and abs(sir - 10 * np.log10(4 / 0.25)) < 1e-9
    assert abs(sar - 10 * np.log10(4 / 0.01)) < 1e-9
    rng = np.random.RandomState(0)
    s, noise = rng.standard_normal(16000), 0.1 * rng.standard_normal(16000)
    m = ev.file_metrics(s.astype(np.float32), (s + noise).astype(np.float32), (s + 0.01 * noise).astype(np.float32))
    assert set(m) == {"pesq", "estoi", "si_sdr", "si_sir", "si_sar"} and m["si_sdr"] > 35
    # wav round trip at 16 bit
    x = (0.5 * np.sin(np.arange(4000) * 0.01)).astype(np.float32)
    ev.write_wav(str(tmp_path / "a.wav"), x)
    assert np.abs(ev.read_wav(str(tmp_path / "a.wav")) - x).max() <= 1.0 / 32768
    # synthetic VoiceBank-DEMAND-shaped set: deterministic, durations inside [1, 10] s
    a, b = ev.synthetic_test_set(6), ev.synthetic_test_set(6)
    assert all(np.array_equal(p[2], q[2]) for p, q in zip(a, b))
    assert all(16000 <= len(p[2]) <= 160000 for p in a)


def test_load_reference_style_checkpoint_with_pickled_data_module_cls(tmp_path, synthetic_sd):
    """a checkpoint as the reference writes it - Lightning layout, torch_ema state under 'ema', and
    hyper_parameters['data_module_cls'] pickled BY REFERENCE as flowmse.data_module.SpecsDataModule - loads without the
    reference package installed; EMA weights become live on eval(), no_ema keeps the raw ones (model.py:92-106)."""
    import sys
    from flowmse_b200 import checkpoint as ck
    from flowmse_b200.model import VFModel, reference_pickle_shims
    ema = {k: v + 0.25 for k, v in synthetic_sd.items()}
    hp = dict(ck.DEFAULT_HPARAMS)
    with reference_pickle_shims():                       # the writer side: the class object the reference would pickle
        import flowmse.data_module as dm
        hp["data_module_cls"] = dm.SpecsDataModule
        c = ck.make_lightning_checkpoint(synthetic_sd, ema, hparams=hp, include_frozen_in_ema=False)
        path = tmp_path / "epoch=7-pesq=2.50.ckpt"
        torch.save(c, path)
    assert "flowmse.data_module" not in sys.modules      # the shim does not outlive the load / save
    model = VFModel.load_from_checkpoint(str(path), base_dir="", batch_size=8, num_workers=4, kwargs=dict(gpu=False))
    assert "flowmse.data_module" not in sys.modules
    name = "all_modules.4.Conv_0.weight"
    assert torch.equal(model.dnn.state_dict()[name], synthetic_sd[name])           # live weights after load
    model.eval(no_ema=False)
    assert torch.equal(model.dnn.state_dict()[name], ema[name])                    # EMA selected by eval()
    assert torch.equal(model.dnn.state_dict()["all_modules.0.W"], synthetic_sd["all_modules.0.W"])   # frozen W: 646 shadows
    model.eval(no_ema=True)
    assert torch.equal(model.dnn.state_dict()[name], synthetic_sd[name])
    assert model.ode.sigma_max == hp.get("sigma_max", 0.487) and model.t_eps == hp.get("t_eps", 0.03)


def test_rk45_restatement_matches_scipy_on_cpu(monkeypatch):
    """The Dormand-Prince tableau, initial-step selection and step-size controller of sampling._rk45_on_device against
    scipy.integrate.solve_ivp(method="RK45") - the solver the reference's get_black_box_solver calls
    (sampling/__init__.py:64-114) - on a small complex system.  The libflowse kernel behind Context.rk_lincomb is replaced
    by its torch definition here (no GPU); the GPU test checks the kernel itself and the device run."""
    import numpy as np
    from scipy import integrate
    import flowmse_b200.runtime as rt
    import flowmse_b200.sampling as S

    class TorchLincomb:
        def rk_lincomb(self, base64, K32, coef, out64=None, out32=None, norm_of=None, rtol=0.0, atol=0.0):
            v = torch.zeros_like(K32[0], dtype=torch.complex128) if base64 is None else base64.clone()
            for s, c in enumerate(coef):
                v = v + c * K32[s].to(torch.complex128)
            if out64 is not None:
                out64.copy_(v)
            if out32 is not None:
                out32.copy_(v.to(torch.complex64))
            if norm_of is not None:
                scale = atol + rtol * torch.maximum(norm_of[0].abs(), norm_of[1].abs())
                return float(((v.abs() / scale) ** 2).sum())

    monkeypatch.setattr(rt, "get_context", lambda dev: TorchLincomb())
    g = torch.Generator().manual_seed(0)
    A = torch.randn(16, 16, generator=g) * 0.8
    y = torch.view_as_complex(torch.randn(1, 1, 4, 4, 2, generator=g))
    x0 = torch.view_as_complex(torch.randn(1, 1, 4, 4, 2, generator=g))

    def VF(x, t, yy):          # stiff towards t -> 0 like the network's 1/t output scale
        v = (x.reshape(-1).real @ A.T) + 1j * (x.reshape(-1).imag @ A.T)
        return (-(v.reshape(x.shape)) * (0.2 / t.reshape(-1, 1, 1, 1)) + 0.1 * yy).to(torch.complex64)

    def ode_func(t, flat):
        xt = torch.from_numpy(flat.reshape(tuple(y.shape))).type(torch.complex64)
        return VF(xt, torch.ones(1) * t, y).numpy().reshape(-1)

    for tol in (1e-3, 1e-6):
        state, nfe = S._rk45_on_device(VF, x0, y, 1.0, 0.03, tol, tol)
        sol = integrate.solve_ivp(ode_func, (1.0, 0.03), x0.numpy().reshape(-1), rtol=tol, atol=tol, method="RK45")
        assert nfe == sol.nfev
        assert np.abs(state.numpy().reshape(-1) - sol.y[:, -1]).max() < 1e-6   # coefficient products differ in the last fp64 bit
