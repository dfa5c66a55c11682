"""Randomised shapes against the fp64 oracle, and the row counts at which the forward switches kernel forms.

The product picks a different kernel form by size: one row tile (K-split clusters, N-split LayerNorm), a few tiles (narrow
column tiles), many tiles (2-SM MMAs), at least as many tile pairs as SM pairs (the stream-K fused FFN with its partial
hand-over).  The fixtures and the full-size cases sit well inside those regimes; these cases are drawn across them and
placed on their boundaries.  Staged protocol of tests/test_gpu_forward.py."""
import numpy as np
import pytest
import torch

from oracle import fs2_oracle as O
from gpu_util import DEV, err_stats, model_for, run
from helpers import OUT_NAMES, call, valid_rows
from test_gpu_forward import TOL_MEL_MAX, TOL_MEL_MEAN, TOL_PRED, check_durations, check_frame_side, check_phoneme_side, log_diag, teacher_kwargs

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 25))
    hi = int(rng.choice([3, 17, 40, 128, 160]))
    lens = rng.integers(1, hi + 1, B).tolist()
    controls = dict(p_control=float(rng.choice([0.7, 1.0, 1.3])), e_control=float(rng.choice([0.5, 1.0])),
                    d_control=float(rng.choice([0.6, 1.0, 1.0, 1.7])))
    return lens, controls


@pytest.mark.parametrize("seed", range(12))
def test_random_batch_against_oracle(seed, sd32, syn):
    lens, controls = _case(seed)
    model = model_for(sd32)
    batch = syn.make_batch(lens, seed=500 + seed)
    want = dict(zip(OUT_NAMES, [t.numpy() if torch.is_tensor(t) else t
                                for t in call(O.forward, batch, O.cast_state_dict(sd32, torch.float64), **controls)]))
    tag = f"fuzz[{seed}] B={len(lens)} P={sum(lens)} F={int(want['mel_lens'].sum())} {controls}"
    free = run(model, batch, **controls)
    # the returned pitch is raw * p_control: so is its error (log_d is control independent)
    mx, _ = err_stats(valid_rows(free[4].cpu().numpy(), lens), valid_rows(want["log_d"], lens))
    assert mx <= TOL_PRED, (tag, "log_d", mx)
    mx, _ = err_stats(valid_rows(free[2].cpu().numpy(), lens), valid_rows(want["pitch"], lens))
    assert mx <= TOL_PRED * max(controls["p_control"], 1.0), (tag, "pitch", mx)
    assert np.array_equal(free[6].cpu().numpy(), want["src_mask"])
    check_durations(tag, free, want, controls["d_control"])
    if int(want["mel_lens"].max()) == 0:
        return
    got = run(model, batch, **teacher_kwargs(want, lens), **controls)
    assert np.array_equal(got[9].cpu().numpy(), want["mel_lens"]) and np.array_equal(got[7].cpu().numpy(), want["mel_mask"])
    check_frame_side(tag, got, want, want["mel_lens"].tolist())


def _frames_case(syn, target_rows, n_utts):
    """n_utts utterances of 40 phonemes whose forced durations add up to `target_rows` packed frame rows (12 reserved rows
    before the first utterance and after every one)."""
    lens = [40] * n_utts
    batch = syn.make_batch(lens, seed=77)
    frames = target_rows - 12 * (n_utts + 1)
    per = frames // n_utts
    d = torch.zeros(n_utts, 40)
    for b in range(n_utts):
        t = per + (frames - per * n_utts if b == 0 else 0)
        base, extra = divmod(t, 40)
        d[b] = base
        d[b, :extra] += 1
    return batch, lens, d


@pytest.mark.parametrize("rows", [128, 129, 18816, 18817, 18945, 19100])
def test_frame_rows_at_the_kernel_form_boundaries(rows, sd32, syn):
    """Packed decoder rows right at one row tile (128 | 129) and at the fused-FFN threshold of 148 row tiles (18,816 | 18,817 rows): below it
    conv9 + w2/LayerNorm, above it the stream-K kernel whose first clusters hand partial tiles over.  Teacher-forced durations
    make the row count exact; the decoder of a slice of the utterances is compared with the fp64 oracle run on that slice."""
    n_utts = 1 if rows < 1000 else 48
    model = model_for(sd32)
    batch, lens, d = _frames_case(syn, rows, n_utts)
    sd64 = O.cast_state_dict(sd32, torch.float64)
    idx = [0] if n_utts == 1 else [0, 23, 47]
    sub = {k: (v[idx] if torch.is_tensor(v) else v) for k, v in batch.items()}
    want = dict(zip(OUT_NAMES, [t.numpy() if torch.is_tensor(t) else t for t in call(O.forward, sub, sd64, d_targets=d[idx].double(),
                                mel_lens=d[idx].sum(1).long(), max_mel_len=int(d.sum(1).max()))]))
    # the whole batch on the GPU with the slice's pitch / energy forced where the slice is (free elsewhere: utterances are independent)
    got = run(model, batch, d_targets=d, mel_lens=d.sum(1).long(), max_mel_len=int(d.sum(1).max()))
    assert model.last_total_frames == rows - 12 * (n_utts + 1)
    p_t, e_t = got[2].clone(), got[3].clone()
    p_t[idx] = torch.as_tensor(want["pitch"]).float().to(DEV)
    e_t[idx] = torch.as_tensor(want["energy"]).float().to(DEV)
    got = run(model, batch, d_targets=d, p_targets=p_t, e_targets=e_t, mel_lens=d.sum(1).long(), max_mel_len=int(d.sum(1).max()))
    T = want["mel_lens"].tolist()
    for i, name in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(got[i][idx].cpu().numpy(), T), valid_rows(want[name], T))
        log_diag(f"boundary rows={rows} {name}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN, (rows, name, mx, mean)
