"""The CLI drop-in end to end on the GPU (random-init weights, mel .npy output)."""
import json
import os
import tempfile

import numpy as np
import pytest
import yaml

import fs2_b200
from fs2_b200 import cli

pytestmark = pytest.mark.gpu


def test_cli_single_sentence(sd32):
    d = tempfile.mkdtemp(prefix="fs2_cli_")
    fs2_b200.synthetic.write_fixture_jsons(d)
    paths = {}
    cfgs = {"p": fs2_b200.config.default_preprocess_config(d), "m": fs2_b200.config.default_model_config(),
            "t": {"path": {"ckpt_path": d, "result_path": os.path.join(d, "result")}}}
    for k, v in cfgs.items():
        paths[k] = os.path.join(d, k + ".yaml")
        with open(paths[k], "w") as f:
            yaml.safe_dump(v, f)
    cli.main(["--restore_step", "1", "--mode", "single", "--text", "今天天气真好", "--speaker_id", "0001", "--emotion",
              "Happy", "-p", paths["p"], "-m", paths["m"], "-t", paths["t"], "--random_init", "--duration_control", "1.2"])
    mel = np.load(os.path.join(d, "result", "synthesis_0001_Happy.npy"))
    meta = json.load(open(os.path.join(d, "result", "synthesis_0001_Happy.json")))
    assert mel.ndim == 2 and mel.shape[1] == 80 and mel.shape[0] == meta["n_frames"] > 16
    # the per-frame pitch / energy curves of synth_samples (utils/tools.py:229-243): phoneme values expanded by the durations
    for key in ("pitch", "energy"):
        curve = np.load(os.path.join(d, "result", f"synthesis_0001_Happy.{key}.npy"))
        assert curve.shape == (meta["n_frames"],) and np.isfinite(curve).all()
    assert meta["n_phonemes"] == 16 and np.isfinite(mel).all()
    # the vocoder step of the reference script (utils/tools.py:258-271): an int16 wav of n_frames * hop samples
    from scipy.io import wavfile
    rate, wav = wavfile.read(os.path.join(d, "result", "synthesis_0001_Happy.wav"))
    assert rate == 22050 and wav.dtype == np.int16 and wav.shape == (meta["n_frames"] * 256,)
    assert np.abs(wav.astype(np.int32)).max() > 100


def test_cli_batch_mode_writes_one_wav_per_line(sd32):
    """--mode batch --source file (dataset_chinese.py:193-276 line format) -> mel + wav per utterance."""
    d = tempfile.mkdtemp(prefix="fs2_cli_")
    fs2_b200.synthetic.write_fixture_jsons(d)
    cfgs = {"p": fs2_b200.config.default_preprocess_config(d), "m": fs2_b200.config.default_model_config(),
            "t": {"path": {"ckpt_path": d, "result_path": os.path.join(d, "result")}}}
    paths = {}
    for k, v in cfgs.items():
        paths[k] = os.path.join(d, k + ".yaml")
        with open(paths[k], "w") as f:
            yaml.safe_dump(v, f)
    src = os.path.join(d, "val.txt")
    with open(src, "w", encoding="utf-8") as f:
        f.write("u1|0001|{j i n t ia n}|今天|Happy|0.8|0.8\n")
        f.write("u2|0003|{n i h ao sh i j ie}|你好世界|Sad|0.3|0.2\n")
        f.write("u3|0002|{h ao}|好|Neutral|0.5|0.5\n")
    cli.main(["--restore_step", "1", "--mode", "batch", "--source", src, "-p", paths["p"], "-m", paths["m"], "-t", paths["t"],
              "--random_init"])
    from scipy.io import wavfile
    for name in ("u1", "u2", "u3"):
        mel = np.load(os.path.join(d, "result", f"{name}.npy"))
        rate, wav = wavfile.read(os.path.join(d, "result", f"{name}.wav"))
        assert wav.shape == (mel.shape[0] * 256,) and wav.dtype == np.int16


def test_host_entry_packed_equals_padded(sd32):
    """synthesize_host: the packed per-utterance views (one D2H of the library's frame-side rows) equal the slices of the
    padded tensor that `synth_samples` takes (utils/tools.py:228-233)."""
    from gpu_util import model_for
    syn = fs2_b200.synthetic
    model = model_for(sd32)
    b = syn.make_batch([30, 7, 19, 30], seed=12)
    host = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in b.items()}
    mels, lens, h2d, d2h_packed = model.synthesize_host(host)
    mels = [np.array(m) for m in mels]
    padded, lens2, _, d2h_padded = model.synthesize_host(host, padded=True)
    assert np.array_equal(lens, lens2) and d2h_packed <= d2h_padded
    for i, n in enumerate(lens):
        assert mels[i].shape == (int(n), 80)
        assert np.array_equal(mels[i], padded[i, : int(n)])


def test_host_entry_reads_what_synth_samples_reads(sd32):
    """utils/tools.py:228-243 reads predictions[2] (pitch), [3] (energy), [5] (durations) and [9] besides the mel: the
    host entry returns them too, equal to the device tensors of a plain forward; results own their memory by default
    (a second call must not change them) and the pinned staging does not grow with the number of distinct batch shapes."""
    from gpu_util import model_for, run
    syn = fs2_b200.synthetic
    model = model_for(sd32)
    b = syn.make_batch([30, 7, 19, 30], seed=12)
    host = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in b.items()}
    mels, lens, _, d2h = model.synthesize_host(host)
    feats = {k: v.copy() for k, v in model.last_host.items()}
    snapshot = [m.copy() for m in mels]
    out = run(model, b)
    assert np.array_equal(feats["pitch"], out[2].cpu().numpy()) and np.array_equal(feats["energy"], out[3].cpu().numpy())
    assert np.array_equal(feats["log_d"], out[4].cpu().numpy()) and np.array_equal(feats["durations"], out[5].cpu().numpy())
    assert np.array_equal(lens, out[9].cpu().numpy())
    assert d2h >= sum(int(n) for n in lens) * 320 + 3 * 4 * 30 * 4
    model.synthesize_host(host)        # (the output staging is double buffered: both slots exist after two calls)
    n_bufs = len(model._pinned)
    for n in (1, 2, 3, 5, 4):          # other batch shapes, then the first again
        other = syn.make_batch([11 + n] * n, seed=n)
        model.synthesize_host({k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in other.items()})
    assert len(model._pinned) == n_bufs, "one staging buffer per role, not per shape"
    for m, s0 in zip(mels, snapshot):
        assert np.array_equal(m, s0), "results of an earlier call must survive later calls"
    views, _, _, _ = model.synthesize_host(host, copy=False)
    assert all(np.array_equal(v, s0) for v, s0 in zip(views, snapshot))


def test_async_host_entry_overlaps_without_mixing_results(sd32):
    """synthesize_host_async: two batches in flight (the second submitted before the first is read) give the same arrays
    as two synchronous calls; views of the first stay intact until its slot is reused."""
    from gpu_util import model_for
    syn = fs2_b200.synthetic
    model = model_for(sd32)
    mk = lambda lens, seed: {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in syn.make_batch(lens, seed=seed).items()}
    a, b = mk([30, 7, 19], 12), mk([11, 40], 13)
    want_a, la, _, _ = model.synthesize_host(a)
    want_b, lb, _, _ = model.synthesize_host(b)
    ha = model.synthesize_host_async(a, copy=False)
    hb = model.synthesize_host_async(b, copy=False)
    got_a, ga_l, _, _ = ha.wait()
    got_b, gb_l, _, _ = hb.wait()
    assert np.array_equal(ga_l, la) and np.array_equal(gb_l, lb)
    assert all(np.array_equal(x, y) for x, y in zip(got_a, want_a)) and all(np.array_equal(x, y) for x, y in zip(got_b, want_b))
