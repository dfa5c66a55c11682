"""Generate tests/golden/*.npz by running the UNMODIFIED reference module.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
It imports the reference with the two import shims of SURVEY.md §8(c) (a fake
matplotlib; `text` registered as a bare namespace so text/__init__.py's missing
third-party imports never run), builds `FastSpeech2(preprocess_config, model_config)`
from the reference's own YAML, loads the synthetic state dict with strict=True (which
pins the 240-key schema), switches to eval + float64 and records inputs and outputs.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, REPO)


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    text = types.ModuleType("text")
    text.__path__ = [os.path.join(REF, "text")]
    sys.modules.setdefault("text", text)
    if REF not in sys.path:
        sys.path.insert(1, REF)
    import model.modules as ref_modules
    import utils.tools as ref_tools
    from model.fastspeech2 import FastSpeech2
    ref_modules.device = torch.device("cpu")
    ref_tools.device = torch.device("cpu")
    return FastSpeech2


def build_reference_model(sd, dtype=torch.float64, pitch_level=None, energy_level=None):
    import fs2_b200
    FastSpeech2 = import_reference()
    cfg_dir = os.path.join(REF, "config", "ESD-Chinese-Singing-MFA")
    preprocess = yaml.load(open(os.path.join(cfg_dir, "preprocess.yaml")), Loader=yaml.FullLoader)
    model_cfg = yaml.load(open(os.path.join(cfg_dir, "model.yaml")), Loader=yaml.FullLoader)
    if pitch_level:      # preprocess.yaml preprocessing.pitch.feature (model/modules.py:28-35)
        preprocess["preprocessing"]["pitch"]["feature"] = pitch_level
    if energy_level:
        preprocess["preprocessing"]["energy"]["feature"] = energy_level
    tmp = tempfile.mkdtemp(prefix="fs2_fixture_")
    fs2_b200.synthetic.write_fixture_jsons(tmp)
    preprocess["path"]["preprocessed_path"] = tmp
    model = FastSpeech2(preprocess, model_cfg)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    return model.to(dtype)


def run_reference(model, batch, **kw):
    with torch.no_grad():
        return model(batch["speakers"], batch["emotions"], batch["arousals"], batch["valences"],
                     batch["texts"], batch["src_lens"], batch["max_src_len"], **kw)


NAMES = ["mel", "postnet", "pitch", "energy", "log_d", "d_rounded", "src_mask", "mel_mask", "src_lens", "mel_lens"]


def golden_cases():
    """(name, batch, forward kwargs, seed of the state dict, mel row stride kept in the fixture)."""
    import fs2_b200
    syn = fs2_b200.synthetic
    cases = []
    cases.append(("c1_single", syn.config1_batch(), {}, 1))
    # pad-row semantics: n_pad = L_max - L_i in {0,1,2,3,15}, mixed conditioning
    cases.append(("pads", syn.make_batch([24, 23, 22, 21, 9, 24], seed=5), {}, 1))
    # control semantics: e_control is dead, energy follows p_control, d_control fractional
    cases.append(("controls", syn.make_batch([17, 12, 20, 5], seed=6),
                  {"p_control": 1.3, "e_control": 0.6, "d_control": 1.5}, 1))
    cases.append(("controls_slow", syn.make_batch([11, 14], seed=7),
                  {"p_control": 0.75, "e_control": 2.0, "d_control": 0.5}, 1))
    return cases


def frame_level_fixtures(out_dir, sd):
    """frame_level pitch / energy: the same state dict under the two other feature
    configurations, free-running and teacher-forced (targets on the frame axis)."""
    import fs2_b200
    syn = fs2_b200.synthetic
    for tag, pl, el in (("frame_both", "frame_level", "frame_level"), ("frame_energy", "phoneme_level", "frame_level"),
                        ("frame_pitch", "frame_level", "phoneme_level")):
        model = build_reference_model(sd, pitch_level=pl, energy_level=el)
        batch = syn.make_batch([14, 9, 14, 3], seed=21)
        kw = {"p_control": 1.2, "d_control": 1.0} if tag == "frame_both" else {}
        out = run_reference(model, batch, **kw)
        rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
        for k, v in kw.items():
            rec[f"kw_{k}"] = np.array(v)
        for n, v in zip(NAMES, out):
            rec[f"out_{n}"] = v.numpy()
        rec["cfg_pitch_level"], rec["cfg_energy_level"] = np.array(pl), np.array(el)
        np.savez_compressed(os.path.join(out_dir, f"{tag}.npz"), **rec)
        print(tag, "mel_lens", out[9].tolist(), "pitch", tuple(out[2].shape), "energy", tuple(out[3].shape))
        if tag != "frame_both":
            continue
        # teacher-forced: targets of frame_level features are [B, max_mel_len]; max_mel_len > max(mel_lens)
        g = torch.Generator().manual_seed(123)
        d_t = torch.randint(0, 7, out[5].shape, generator=g) * (~out[6])
        mel_lens = d_t.sum(1)
        T = int(mel_lens.max()) + 5
        fmask = torch.arange(T).unsqueeze(0) >= mel_lens.unsqueeze(1)
        p_t = (torch.randn(d_t.shape[0], T, generator=g, dtype=torch.float64) * 1.5) * (~fmask)
        e_t = (torch.randn(d_t.shape[0], T, generator=g, dtype=torch.float64) * 1.5) * (~fmask)
        kw = dict(mel_lens=mel_lens, max_mel_len=T, p_targets=p_t, e_targets=e_t, d_targets=d_t)
        out = run_reference(model, batch, **kw)
        rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
        for k, v in kw.items():
            rec[f"kw_{k}"] = v.numpy() if torch.is_tensor(v) else np.array(v)
        for n, v in zip(NAMES, out):
            rec[f"out_{n}"] = v.numpy()
        rec["cfg_pitch_level"], rec["cfg_energy_level"] = np.array(pl), np.array(el)
        np.savez_compressed(os.path.join(out_dir, "frame_both_forced.npz"), **rec)
        print("frame_both_forced mel_lens", out[9].tolist())


LOG_STATS = {"pitch": [0.35, 9.0, 2.0, 1.0], "energy": [0.2, 8.0, 2.0, 1.0]}    # positive minima: log bins need them


def log_quantisation_fixture(out_dir, sd):
    """variance_embedding.{pitch,energy}_quantization = "log" (model/modules.py:48-54,60-66): bins =
    exp(linspace(log(min), log(max), n_bins - 1)) from a stats.json with positive minima.  The predictor head biases are
    raised so that the predictions land inside the bins.  The bins the reference constructs are recorded (the facade
    must build the same ones from the same configuration) and loaded back through the state dict as usual."""
    import json
    import fs2_b200
    FastSpeech2 = import_reference()
    syn = fs2_b200.synthetic
    cfg_dir = os.path.join(REF, "config", "ESD-Chinese-Singing-MFA")
    preprocess = yaml.load(open(os.path.join(cfg_dir, "preprocess.yaml")), Loader=yaml.FullLoader)
    model_cfg = yaml.load(open(os.path.join(cfg_dir, "model.yaml")), Loader=yaml.FullLoader)
    model_cfg["variance_embedding"]["pitch_quantization"] = "log"
    model_cfg["variance_embedding"]["energy_quantization"] = "log"
    tmp = tempfile.mkdtemp(prefix="fs2_fixture_")
    syn.write_fixture_jsons(tmp)
    with open(os.path.join(tmp, "stats.json"), "w") as f:
        json.dump(LOG_STATS, f)
    preprocess["path"]["preprocessed_path"] = tmp
    model = FastSpeech2(preprocess, model_cfg)
    bins = {k: model.state_dict()[f"variance_adaptor.{k}_bins"].clone() for k in ("pitch", "energy")}
    sd = dict(sd)
    for k in ("pitch", "energy"):
        sd[f"variance_adaptor.{k}_bins"] = bins[k]
        sd[f"variance_adaptor.{k}_predictor.linear_layer.bias"] = torch.tensor([2.0])
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(torch.float64)
    batch = syn.make_batch([19, 8, 23, 15], seed=33)
    kw = {"p_control": 1.25}
    out = run_reference(model, batch, **kw)
    rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
    rec["kw_p_control"] = np.array(kw["p_control"])
    for n, v in zip(NAMES, out):
        rec[f"out_{n}"] = v.numpy()
    for k in ("pitch", "energy"):
        rec[f"ref_{k}_bins"] = bins[k].numpy()
        rec[f"cfg_stats_{k}"] = np.array(LOG_STATS[k])
        idx = torch.bucketize(out[2 if k == "pitch" else 3], bins[k].double())
        print("log_bins", k, "bucket range", int(idx.min()), int(idx.max()), "distinct", len(torch.unique(idx)))
    np.savez_compressed(os.path.join(out_dir, "log_bins.npz"), **rec)
    print("log_bins mel_lens", out[9].tolist())


def main():
    if "--log-bins-only" in sys.argv:
        import fs2_b200
        log_quantisation_fixture(os.path.dirname(os.path.abspath(__file__)), fs2_b200.synthetic.synthetic_state_dict(seed=0))
        return
    if "--frame-level-only" in sys.argv:    # adds the frame_level fixtures without rewriting the others
        import fs2_b200
        frame_level_fixtures(os.path.dirname(os.path.abspath(__file__)), fs2_b200.synthetic.synthetic_state_dict(seed=0))
        return
    import fs2_b200
    syn = fs2_b200.synthetic
    out_dir = os.path.dirname(os.path.abspath(__file__))
    sd = syn.synthetic_state_dict(seed=0)
    model = build_reference_model(sd)
    n_keys = len(model.state_dict())
    print("reference state-dict keys:", n_keys, "params:", sum(p.numel() for p in model.parameters()))
    meta = {"checksum": syn.state_dict_checksum(sd), "n_keys": n_keys}
    np.savez(os.path.join(out_dir, "meta.npz"), **{k: np.array(v) for k, v in meta.items()})
    with open(os.path.join(out_dir, "state_dict_keys.txt"), "w") as f:
        for k, v in build_reference_model(sd, torch.float32).state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {str(v.dtype).replace('torch.', '')}\n")

    for name, batch, kw, stride in golden_cases():
        out = run_reference(model, batch, **kw)
        rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
        for k, v in kw.items():
            rec[f"kw_{k}"] = np.array(v)
        for n, v in zip(NAMES, out):
            rec[f"out_{n}"] = v.numpy()
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **rec)
        print(name, "mel_lens", out[9].tolist())

    # teacher-forced case: durations / pitch / energy / mel_lens / max_mel_len supplied
    # (model/fastspeech2.py:82-87), with max_mel_len > max(mel_lens) so T_max tails differ.
    batch = syn.make_batch([13, 9, 16], seed=8)
    free = run_reference(model, batch)
    g = torch.Generator().manual_seed(99)
    d_t = torch.randint(0, 9, free[5].shape, generator=g) * (~free[6])
    p_t = (torch.randn(free[2].shape, generator=g, dtype=torch.float64) * 1.5) * (~free[6])
    e_t = (torch.randn(free[3].shape, generator=g, dtype=torch.float64) * 1.5) * (~free[6])
    mel_lens = d_t.sum(1)
    kw = dict(mel_lens=mel_lens, max_mel_len=int(mel_lens.max()) + 7, p_targets=p_t, e_targets=e_t, d_targets=d_t)
    out = run_reference(model, batch, **kw)
    rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
    for k, v in kw.items():
        rec[f"kw_{k}"] = v.numpy() if torch.is_tensor(v) else np.array(v)
    for n, v in zip(NAMES, out):
        rec[f"out_{n}"] = v.numpy()
    np.savez_compressed(os.path.join(out_dir, "teacher_forced.npz"), **rec)
    print("teacher_forced mel_lens", out[9].tolist())

    # long-form: crosses max_seq_len=2000 on the decoder side (Models.py:145-152); keep every 13th frame
    batch = syn.make_batch([400], seed=9)
    out = run_reference(model, batch, d_control=1.0)
    rec = {f"in_{k}": (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in batch.items()}
    rec["kw_d_control"] = np.array(1.0)
    for n, v in zip(NAMES, out):
        a = v.numpy()
        rec[f"out_{n}"] = a[:, ::13] if n in ("mel", "postnet") else a
    rec["mel_row_stride"] = np.array(13)
    np.savez_compressed(os.path.join(out_dir, "longform.npz"), **rec)
    print("longform mel_lens", out[9].tolist())
    frame_level_fixtures(out_dir, sd)
    log_quantisation_fixture(out_dir, sd)


if __name__ == "__main__":
    main()
