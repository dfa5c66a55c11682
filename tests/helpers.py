"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT_NAMES = ["mel", "postnet", "pitch", "energy", "log_d", "d_rounded", "src_mask", "mel_mask", "src_lens", "mel_lens"]


def golden_names(frame_level=None):
    """All fixtures; frame_level=False -> only the default (phoneme_level) configuration, True -> only the
    fixtures recorded with a frame_level pitch and/or energy feature."""
    names = sorted(os.path.splitext(os.path.basename(p))[0]
                   for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                   if os.path.basename(p) not in ("meta.npz", "vocoder.npz", "log_bins.npz"))
    if frame_level is None:
        return names
    return [n for n in names if n.startswith("frame_") == frame_level]


def golden_levels(name):
    """preprocessing.{pitch,energy}.feature the fixture was recorded with (model/modules.py:28-35)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    if "cfg_pitch_level" not in z.files:
        return {"pitch_level": "phoneme_level", "energy_level": "phoneme_level"}
    return {"pitch_level": str(z["cfg_pitch_level"]), "energy_level": str(z["cfg_energy_level"])}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    batch, kw, out = {}, {}, {}
    for k in z.files:
        a = z[k]
        if k.startswith("in_"):
            batch[k[3:]] = int(a) if a.ndim == 0 else torch.from_numpy(a)
        elif k.startswith("kw_"):
            n = k[3:]
            if a.ndim == 0:
                kw[n] = int(a) if n == "max_mel_len" else float(a)
            else:
                kw[n] = torch.from_numpy(a)
        elif k.startswith("out_"):
            out[k[4:]] = a
    stride = int(z["mel_row_stride"]) if "mel_row_stride" in z.files else 1
    return batch, kw, out, stride


def call(fn, batch, *lead, **kw):
    return fn(*lead, batch["speakers"], batch["emotions"], batch["arousals"], batch["valences"],
              batch["texts"], batch["src_lens"], batch["max_src_len"], **kw)


def valid_rows(arr, lens):
    """Concatenate arr[b, :lens[b]] over the batch (padding rows are outside the contract)."""
    return np.concatenate([np.asarray(arr[b, : int(lens[b])]) for b in range(len(lens))], axis=0)


def log_bins_case(sd):
    """The `log` quantisation fixture (model/modules.py:48-54,60-66; tests/golden/make_golden.py:log_quantisation_fixture):
    returns (batch, kwargs, reference outputs, state dict with the fixture's head biases and the REFERENCE's bins,
    stats.json content the bins were built from)."""
    z = np.load(os.path.join(GOLDEN_DIR, "log_bins.npz"))
    batch, kw, out, _ = load_golden("log_bins")
    sd = dict(sd)
    stats = {}
    for k in ("pitch", "energy"):
        sd[f"variance_adaptor.{k}_bins"] = torch.from_numpy(z[f"ref_{k}_bins"]).to(sd[f"variance_adaptor.{k}_bins"].dtype)
        sd[f"variance_adaptor.{k}_predictor.linear_layer.bias"] = torch.tensor([2.0], dtype=sd["mel_linear.bias"].dtype)
        stats[k] = [float(v) for v in z[f"cfg_stats_{k}"]]
    return batch, kw, out, sd, stats


def log_bins_model(stats, math_mode="tf32"):
    """FastSpeech2B200 constructed from a `log` model config and the fixture's stats.json: its bins come from the
    facade's own construction (model.py), not from a state dict."""
    import json
    import tempfile
    import fs2_b200
    d = fs2_b200.synthetic.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    with open(os.path.join(d, "stats.json"), "w") as f:
        json.dump(stats, f)
    cfg = fs2_b200.config.default_model_config()
    cfg["variance_embedding"]["pitch_quantization"] = "log"
    cfg["variance_embedding"]["energy_quantization"] = "log"
    return fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(d), cfg, math_mode=math_mode)
