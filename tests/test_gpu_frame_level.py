"""frame_level pitch / energy features (preprocess.yaml preprocessing.{pitch,energy}.feature; model/modules.py:28-35,
139-148): the predictors run on the expanded frame rows after the LengthRegulator and the
predictions come back as [B, max_mel_len].  Checked against fixtures recorded from the unmodified reference
(tests/golden/frame_*.npz) with the staged protocol of tests/test_gpu_forward.py; same TF32 tolerances."""
import numpy as np
import pytest
import torch

from gpu_util import err_stats, model_for, run
from helpers import golden_levels, golden_names, load_golden, valid_rows
from test_gpu_forward import TOL_PRED, check_durations, check_frame_side, log_diag

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("math_mode", ["tf32", "bf16"])
@pytest.mark.parametrize("name", golden_names(frame_level=True))
def test_frame_level_fixture(name, math_mode, sd32):
    lv = golden_levels(name)
    model = model_for(sd32, math_mode=math_mode, **lv)
    batch, kw, want, _ = load_golden(name)
    src_lens, mel_lens = batch["src_lens"].tolist(), want["mel_lens"].tolist()
    tol_pred, tol_mel, tol_mean = (TOL_PRED, 1.5e-3, 3e-4) if math_mode == "tf32" else (1.5e-2, 1.2e-2, 2.2e-3)
    tag = f"{name}[{math_mode}]"
    p_frame, e_frame = lv["pitch_level"] == "frame_level", lv["energy_level"] == "frame_level"
    T = int(want["mel"].shape[1])
    f32 = lambda a: torch.as_tensor(a).float()

    def cmp(got, key, lens, what):
        mx, mean = err_stats(valid_rows(got.cpu().numpy(), lens), valid_rows(want[key], lens))
        log_diag(f"{tag} {what}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= tol_pred, (tag, what, mx)

    if "d_targets" in kw:   # everything forced, targets on the frame axis; predictions are returned unscaled
        got = run(model, batch, **{k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in kw.items()})
        assert np.array_equal(got[9].cpu().numpy(), want["mel_lens"])
        assert np.array_equal(got[7].cpu().numpy(), want["mel_mask"])
        assert tuple(got[2].shape) == want["pitch"].shape and tuple(got[3].shape) == want["energy"].shape
        cmp(got[2], "pitch", mel_lens, "pitch (frames, forced)")
        cmp(got[3], "energy", mel_lens, "energy (frames, forced)")
        for i in (2, 3):
            assert (got[i].cpu().numpy()[want["mel_mask"]] == 0).all()
        for i, n in ((0, "mel"), (1, "postnet")):
            mx, mean = err_stats(valid_rows(got[i].cpu().numpy(), mel_lens), valid_rows(want[n], mel_lens))
            log_diag(f"{tag} {n}: max {mx:.3e} mean {mean:.3e}")
            assert mx <= tol_mel and mean <= tol_mean
        return

    # stage A: free-running phoneme side
    free = run(model, batch, **kw)
    cmp(free[4], "log_d", src_lens, "log_d")
    if math_mode == "tf32":
        check_durations(tag, free, want, kw.get("d_control", 1.0))
    # stage B: durations forced -> the frame axis is the reference's.  The features are then checked in the order the
    # reference evaluates them (phoneme_level pitch, energy, then frame_level pitch, energy: modules.py:114-148), each
    # with every feature upstream of it forced to the reference's values (their bucket indices feed the next predictor).
    forced = dict(d_targets=f32(want["d_rounded"]), mel_lens=torch.as_tensor(want["mel_lens"]), max_mel_len=T, **kw)
    order = [f for f in (("pitch", p_frame), ("energy", e_frame)) if not f[1]] + \
            [f for f in (("pitch", p_frame), ("energy", e_frame)) if f[1]]
    for feat, on_frames in order:
        got = run(model, batch, **forced)
        assert np.array_equal(got[9].cpu().numpy(), want["mel_lens"])
        assert tuple(got[2].shape) == want["pitch"].shape and tuple(got[3].shape) == want["energy"].shape
        idx = 2 if feat == "pitch" else 3
        cmp(got[idx], feat, mel_lens if on_frames else src_lens, f"{feat} ({'frames' if on_frames else 'phonemes'})")
        pad = want["mel_mask"] if on_frames else want["src_mask"]
        assert (got[idx].cpu().numpy()[pad] == 0).all()
        forced["p_targets" if feat == "pitch" else "e_targets"] = f32(want[feat])
    # stage C: everything forced -> mel
    got = run(model, batch, **forced)
    assert np.array_equal(got[7].cpu().numpy(), want["mel_mask"])
    for i, n in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(got[i].cpu().numpy(), mel_lens), valid_rows(want[n], mel_lens))
        log_diag(f"{tag} {n}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= tol_mel and mean <= tol_mean


def test_frame_level_target_shape_is_checked(sd32, syn):
    model = model_for(sd32, pitch_level="frame_level", energy_level="frame_level")
    batch = syn.make_batch([8, 6], seed=1)
    with pytest.raises(RuntimeError, match="frame_level"):
        run(model, batch, p_targets=torch.zeros(2, 8))
