"""The evaluate.py drop-in (flowmse_b200/evaluate.py) end to end on a small synthetic test set:
wav files on disk -> bucketed device STFT / sampler / iSTFT -> the reference's output files."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_evaluate_driver_writes_reference_outputs(tmp_path, synthetic_sd):
    from flowmse_b200 import evaluate as ev
    from flowmse_b200.model import VFModel
    # a VoiceBank-DEMAND-shaped directory: <test_dir>/test/{clean,noisy}/*.wav
    root = tmp_path / "data"
    (root / "test" / "clean").mkdir(parents=True)
    (root / "test" / "noisy").mkdir(parents=True)
    pairs = ev.synthetic_test_set(4, seed=3)
    for name, clean, noisy in pairs:
        ev.write_wav(str(root / "test" / "clean" / name), clean * 0.5)
        ev.write_wav(str(root / "test" / "noisy" / name), noisy * 0.5)
    out = tmp_path / "out"
    summary = ev.main(["--test_dir", str(root), "--folder_destination", str(out), "--synthetic_weights", "0", "--N", "2",
                       "--seed", "11", "--max_batch_frames", "1024"])
    assert summary["files"] == 4 and summary["frames"] > 0 and summary["frames_per_s"] > 0
    for f in ("_results.csv", "_avg_results.txt", "_settings.txt", "_timing.json"):
        assert (out / f).exists(), f
    import pandas as pd
    df = pd.read_csv(out / "_results.csv")
    assert list(df.columns) == ["filename", "pesq", "estoi", "si_sdr", "si_sir", "si_sar"]
    assert sorted(df["filename"]) == sorted(p[0] for p in pairs) and np.isfinite(df["si_sdr"]).all()
    settings = (out / "_settings.txt").read_text()
    assert "odesolver: euler" in settings and "N: 2" in settings and "sigma_max: 0.487" in settings
    # every enhanced file has the length of its input and equals the per-file reference flow (VFModel.enhance) when the
    # prior noise is the same: re-run one file alone with the batch's noise reproduced is covered in test_gpu_stft; here
    # check length, finiteness and that the output is not the input
    for name, clean, noisy in pairs:
        xh = ev.read_wav(str(out / "files" / name))
        assert xh.shape == noisy.shape and np.isfinite(xh).all()
        assert np.abs(xh - noisy * 0.5).max() > 1e-3
    assert json.load(open(out / "_timing.json"))["n_gpus"] == 1
