"""BASELINE.json's full-size configurations on the GPU, checked through size-independent
properties (the oracle needs ~10 s per batch-64 forward on a CPU, so only a slice of each case is
compared with it) plus the long-form case against the oracle end to end."""
import numpy as np
import pytest
import torch

from oracle import fs2_oracle as O
from gpu_util import DEV, err_stats, model_for, run
from helpers import OUT_NAMES, call, valid_rows
from test_gpu_forward import TOL_MEL_MAX, TOL_MEL_MEAN, TOL_PRED, check_durations, check_frame_side, check_phoneme_side, log_diag, teacher_kwargs

pytestmark = pytest.mark.gpu


def test_config2_batch64_properties(sd32, syn):
    model = model_for(sd32)
    batch = syn.config2_batch(seed=0)
    out = run(model, batch)
    mel, post, pitch, energy, log_d, d_round, src_mask, mel_mask, src_lens, mel_lens = out
    B, L = batch["texts"].shape
    # shapes / dtypes of the reference tuple
    T = int(mel_lens.max())
    assert mel.shape == (B, T, 80) and post.shape == (B, T, 80) and mel.dtype == torch.float32
    assert pitch.shape == (B, L) and d_round.shape == (B, L) and src_mask.dtype == torch.bool and mel_lens.dtype == torch.int64
    assert all(torch.isfinite(t).all() for t in (mel, post, pitch, energy, log_d, d_round))
    # integer bookkeeping: mel_lens = sum trunc(d), masks from lengths, d = clamp(round(exp(logd)-1)) on the device values
    assert torch.equal(mel_lens, torch.clamp(torch.trunc(d_round), min=0).sum(1).long())
    assert torch.equal(d_round, torch.clamp(torch.round(torch.exp(log_d) - 1) * 1.0, min=0) * (~src_mask))
    assert torch.equal(mel_mask, torch.arange(T, device=DEV)[None, :] >= mel_lens[:, None])
    assert torch.equal(src_mask, torch.arange(L, device=DEV)[None, :] >= batch["src_lens"].to(DEV)[:, None])
    bias = sd32["mel_linear.bias"].to(DEV)
    assert all(torch.equal(mel[b, int(t):], bias.expand(T - int(t), 80)) for b, t in enumerate(mel_lens) if int(t) < T)
    # permutation of the utterances permutes the results bit for bit (teacher-forced: same durations / T_max)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    pb = {k: (v[perm] if torch.is_tensor(v) else v) for k, v in batch.items()}
    forced = dict(d_targets=d_round.cpu(), p_targets=pitch.cpu(), e_targets=energy.cpu())
    # (with the two-launch FFN: every row then sees the same summation order wherever it sits in the packed layout;
    # the stream-K schedule of the fused FFN adds a row tile's hidden chunks in an order that depends on which cluster
    # owns which chunk, i.e. on the tile's position, so under it the permuted results agree to rounding only)
    from gpu_util import lib
    try:
        lib().fs2_debug_set_flag(4, 0)
        a = run(model, batch, **forced)
        b = run(model, pb, **{k: v[perm] for k, v in forced.items()})
    finally:
        lib().fs2_debug_set_flag(4, 2)
    assert torch.equal(a[1][perm], b[1]) and torch.equal(a[0][perm], b[0]) and torch.equal(a[9][perm], b[9])
    a2 = run(model, batch, **forced)
    b2 = run(model, pb, **{k: v[perm] for k, v in forced.items()})
    assert torch.equal(a2[9][perm], b2[9])
    d_fused = max(float((a2[i][perm] - b2[i]).abs().max()) for i in (0, 1))
    d_forms = max(float((a2[i] - a[i]).abs().max()) for i in (0, 1))
    log_diag(f"config2 x64 fused FFN (stream-K): permuted vs not {d_fused:.3e}; fused vs two-launch {d_forms:.3e}")
    # neither is a summation-order effect alone: the fused kernel rounds the hidden activations to TF32 to nearest where the
    # two-launch form lets the tensor core truncate them, and a last-bit difference in one block's output flips TF32
    # operand roundings downstream (one flip = one TF32 ulp); both stay inside the tolerance against the oracle
    assert d_fused <= TOL_MEL_MAX and d_forms <= TOL_MEL_MAX
    a3 = run(model, batch, **forced)     # the hand-over between clusters is deterministic: same layout, same bits
    assert torch.equal(a3[0], a2[0]) and torch.equal(a3[1], a2[1])
    a = a2
    # a slice of the batch against the oracle: 4 utterances alone with the batch's L_max and T_max forced
    idx = [0, 17, 40, 63]
    sub = {k: (v[idx] if torch.is_tensor(v) else v) for k, v in batch.items()}
    want = dict(zip(OUT_NAMES, call(O.forward, sub, O.cast_state_dict(sd32, torch.float64), d_targets=d_round.cpu()[idx].double(),
                                    p_targets=pitch.cpu()[idx].double(), e_targets=energy.cpu()[idx].double(),
                                    mel_lens=mel_lens.cpu()[idx], max_mel_len=T)))
    lens = mel_lens.cpu()[idx].tolist()
    for i, n in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(a[i][idx].cpu().numpy(), lens), valid_rows(want[n].numpy(), lens))
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN, (n, mx, mean)


def test_config2_all_64_utterances_against_oracle(sd32, syn):
    """The WHOLE config-2 batch (64 utterances, 4,134 phonemes, ~26 k frames) against the fp64 oracle under the staged
    protocol: free-running phoneme side (log-duration, pitch; energy with the oracle's pitch forced) for every utterance,
    integer durations exact up to reported rounding boundaries, then the frame side of all 64 teacher-forced."""
    model = model_for(sd32)
    batch = syn.config2_batch(seed=0)
    want = dict(zip(OUT_NAMES, [t.numpy() if torch.is_tensor(t) else t
                                for t in call(O.forward, batch, O.cast_state_dict(sd32, torch.float64))]))
    src_lens = batch["src_lens"].tolist()
    free = run(model, batch)
    check_phoneme_side("config2 x64", free, want, src_lens)
    check_durations("config2 x64", free, want, 1.0)
    forced_p = run(model, batch, p_targets=torch.as_tensor(want["pitch"]).float())
    mx, mean = err_stats(valid_rows(forced_p[3].cpu().numpy(), src_lens), valid_rows(want["energy"], src_lens))
    log_diag(f"config2 x64 energy (pitch forced): max {mx:.3e} mean {mean:.3e}")
    assert mx <= TOL_PRED
    got = run(model, batch, **teacher_kwargs(want, src_lens))
    assert np.array_equal(got[9].cpu().numpy(), want["mel_lens"]) and np.array_equal(got[7].cpu().numpy(), want["mel_mask"])
    check_frame_side("config2 x64", got, want, want["mel_lens"].tolist())


@pytest.mark.parametrize("d_control", [0.5, 2.0])
def test_config4_longform(d_control, sd32, syn):
    """400 phonemes, ~1.3k / ~5.4k frames: crosses max_seq_len=2000 on the decoder side (regenerated
    sinusoid rows, transformer/Models.py:145-152), long attention, large expansion."""
    model = model_for(sd32)
    batch = syn.config4_batch()
    want = dict(zip(OUT_NAMES, call(O.forward, batch, sd32, d_control=d_control)))   # fp32 oracle: ~5 s at 5.4k frames
    free = run(model, batch, d_control=d_control)
    n = batch["src_lens"].tolist()
    mx, _ = err_stats(valid_rows(free[4].cpu().numpy(), n), valid_rows(want["log_d"].numpy(), n))
    assert mx <= TOL_PRED
    got = run(model, batch, d_targets=want["d_rounded"], p_targets=want["pitch"], e_targets=want["energy"],
              mel_lens=want["mel_lens"], max_mel_len=int(want["mel_lens"].max()))
    assert torch.equal(got[9].cpu(), want["mel_lens"])
    T = want["mel_lens"].tolist()
    assert (d_control == 2.0) == (T[0] > 2000)
    for i, name in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(got[i].cpu().numpy(), T), valid_rows(want[name].numpy(), T))
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN, (name, mx, mean)


def test_sharded_batch_matches_oracle_per_shard(sd32, syn):
    """Multi-GPU semantics on one GPU: each LPT shard is its own batch (its own L_max / T_max)."""
    from fs2_b200 import partition
    model = model_for(sd32)
    batch = syn.make_batch(syn.random_lengths(10, lo=10, hi=60, seed=3), seed=61)
    sd64 = O.cast_state_dict(sd32, torch.float64)
    for shard in partition.lpt_partition(batch["src_lens"].tolist(), 2):
        sub = partition.take(batch, shard)
        want = dict(zip(OUT_NAMES, call(O.forward, sub, sd64)))
        got = run(model, sub, d_targets=want["d_rounded"].float(), p_targets=want["pitch"].float(),
                  e_targets=want["energy"].float(), mel_lens=want["mel_lens"], max_mel_len=int(want["mel_lens"].max()))
        T = want["mel_lens"].tolist()
        mx, mean = err_stats(valid_rows(got[1].cpu().numpy(), T), valid_rows(want["postnet"].numpy(), T))
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN


def test_config5_control_sweep_batch32(sd32, syn):
    """BASELINE config 5: pitch / energy / duration controls over 0.5 .. 2.0 on a batch of 32.
    Size-independent properties on the GPU results of the full sweep, plus two corners against the fp32 oracle:
      * e_control never changes anything; energy follows p_control (model/modules.py:123-125);
      * log-durations are control independent; d_rounded = clamp(round(exp(logd) - 1) * d_control) exactly;
      * pitch scales linearly with p_control while the bucket embedding changes (so energy and mel do change);
      * mel_lens = sum trunc(d_rounded) for every d_control."""
    model = model_for(sd32)
    batch = syn.make_batch(syn.random_lengths(32, seed=5), seed=55)
    base = run(model, batch)
    pad = base[6]
    for d in (0.5, 0.75, 1.0, 1.5, 2.0):
        out = run(model, batch, d_control=d, e_control=0.5 + d)
        assert torch.equal(out[4], base[4]) and torch.equal(out[2], base[2]) and torch.equal(out[3], base[3])
        want_d = torch.clamp(torch.round(torch.exp(out[4]) - 1) * d, min=0) * (~pad)
        assert torch.equal(out[5], want_d)
        assert torch.equal(out[9], torch.clamp(torch.trunc(out[5]), min=0).sum(1).long())
        assert torch.isfinite(out[1]).all()
    raw_pitch = base[2]
    for p in (0.5, 0.75, 1.5, 2.0):
        out = run(model, batch, p_control=p, e_control=2.5 - p)
        ref = run(model, batch, p_control=p)
        for x, y in zip(out, ref):
            assert torch.equal(x, y), "e_control changed the output"
        assert torch.allclose(out[2], raw_pitch * p, rtol=0, atol=1e-6)      # pitch = raw * p_control (pads stay 0)
        assert torch.equal(out[4], base[4])                                   # durations are computed before the pitch add
        assert not torch.equal(out[3], base[3])                               # energy sees the new pitch buckets and p_control
    want = call(O.forward, batch, sd32, p_control=2.0, d_control=0.5)
    got = run(model, batch, p_control=2.0, d_control=0.5)
    n = batch["src_lens"].tolist()
    for i in (2, 4):
        mx, _ = err_stats(valid_rows(got[i].cpu().numpy(), n), valid_rows(want[i].numpy(), n))
        assert mx <= TOL_PRED * (2.0 if i == 2 else 1.0)     # the returned pitch is raw * p_control: so is its error


def test_config3_batch512_properties(sd32, syn):
    """BASELINE config 3 on one GPU: 512 utterances, ~200 k frames.  At this size the forward selects the fused FFN kernel
    by itself (>= 4 waves of row tiles), so this is also its end-to-end check: integer bookkeeping, finiteness, and a slice
    of the batch against the fp64 oracle with the batch's L_max / T_max forced."""
    model = model_for(sd32)
    batch = syn.config2_batch(seed=3, batch=512)
    out = run(model, batch)
    mel, post, pitch, energy, log_d, d_round, src_mask, mel_mask, src_lens, mel_lens = out
    T = int(mel_lens.max())
    assert int(mel_lens.sum()) > 150_000 and mel.shape == (512, T, 80)
    assert all(torch.isfinite(t).all() for t in (mel, post, pitch, energy, log_d))
    assert torch.equal(mel_lens, torch.clamp(torch.trunc(d_round), min=0).sum(1).long())
    assert torch.equal(mel_mask, torch.arange(T, device=DEV)[None, :] >= mel_lens[:, None])
    forced = dict(d_targets=d_round.cpu(), p_targets=pitch.cpu(), e_targets=energy.cpu())
    a = run(model, batch, **forced)
    idx = [3, 47, 100, 163, 250, 301, 377, 420, 468, 511]
    sub = {k: (v[idx] if torch.is_tensor(v) else v) for k, v in batch.items()}
    want = dict(zip(OUT_NAMES, call(O.forward, sub, O.cast_state_dict(sd32, torch.float64), d_targets=d_round.cpu()[idx].double(),
                                    p_targets=pitch.cpu()[idx].double(), e_targets=energy.cpu()[idx].double(),
                                    mel_lens=mel_lens.cpu()[idx], max_mel_len=T)))
    lens = mel_lens.cpu()[idx].tolist()
    for i, n in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(a[i][idx].cpu().numpy(), lens), valid_rows(want[n].numpy(), lens))
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN, (n, mx, mean)


def test_export_import_of_stage1_is_bit_exact_and_matches_oracle_after_regrouping(sd32, syn):
    """fs2_export_stage1 / fs2_import_stage1 (the hand-over partition.rebalanced_forward uses to balance a sharded batch by
    FRAMES): (1) decoding the exported rows on the same batch equals the plain forward bit for bit; (2) utterances encoded
    in two phoneme shards and REGROUPED into two other batches decode to the oracle's result on the receiving batch with
    the durations / pitch / energy of stage 1 forced."""
    from fs2_b200 import partition
    model = model_for(sd32)
    batch = syn.make_batch(syn.random_lengths(10, lo=10, hi=60, seed=13), seed=71)
    names = ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")
    plain = run(model, batch)
    b = {k: v.to(DEV) for k, v in batch.items() if torch.is_tensor(v)}
    s1 = model.encode(*[b[k] for k in names], batch["max_src_len"])
    mel, post, mask, lens = model.decode(s1["hidden"], s1["reps"], b["src_lens"])
    torch.cuda.synchronize()
    assert torch.equal(lens, plain[9]) and torch.equal(mel, plain[0]) and torch.equal(post, plain[1]) and torch.equal(mask, plain[7])
    assert torch.equal(s1["pitch"], plain[2]) and torch.equal(s1["energy"], plain[3]) and torch.equal(s1["d_round"], plain[5])
    # (2) two stage-1 shards -> regrouped stage-2 shards
    parts = partition.lpt_partition(batch["src_lens"].tolist(), 2)
    enc = []
    for p in parts:
        sub = partition.take(batch, p)
        sb = {k: v.to(DEV) for k, v in sub.items() if torch.is_tensor(v)}
        enc.append((p, sub, model.encode(*[sb[k] for k in names], sub["max_src_len"])))
        enc[-1][2]["hidden"] = enc[-1][2]["hidden"].clone()      # (the same context encodes both shards here)
    mel_lens_all = [0] * 10
    for p, _, e in enc:
        for j, g in enumerate(p):
            mel_lens_all[g] = int(e["mel_lens"][j])
    new_parts, moves = partition.plan_rebalance(parts, batch["src_lens"].tolist(), mel_lens_all, 2)
    sd64 = O.cast_state_dict(sd32, torch.float64)
    where = {g: (r, j) for r, (p, _, _) in enumerate(enc) for j, g in enumerate(p)}
    for new in new_parts:
        L = max(int(batch["src_lens"][g]) for g in new)
        hid = torch.zeros(len(new), L, 256, device=DEV)
        rep = torch.zeros(len(new), L, dtype=torch.int32, device=DEV)
        d_t, p_t, e_t = (torch.zeros(len(new), L, dtype=torch.float64) for _ in range(3))
        for i, g in enumerate(new):
            r, j = where[g]
            n = int(batch["src_lens"][g])
            e = enc[r][2]
            hid[i, :n], rep[i, :n] = e["hidden"][j, :n], e["reps"][j, :n]
            d_t[i, :n], p_t[i, :n], e_t[i, :n] = e["d_round"][j, :n].cpu(), e["pitch"][j, :n].cpu(), e["energy"][j, :n].cpu()
        sub = partition.take(batch, new)
        got = model.decode(hid, rep, sub["src_lens"].to(DEV))
        want = dict(zip(OUT_NAMES, call(O.forward, sub, sd64, d_targets=d_t, p_targets=p_t, e_targets=e_t,
                                        mel_lens=got[3].cpu(), max_mel_len=int(got[3].max()))))
        assert torch.equal(got[3].cpu(), want["mel_lens"])
        T = want["mel_lens"].tolist()
        for i, n in ((0, "mel"), (1, "postnet")):
            mx, mean = err_stats(valid_rows(got[i].cpu().numpy(), T), valid_rows(want[n].numpy(), T))
            log_diag(f"regrouped stage 2 {n}: max {mx:.3e} mean {mean:.3e}")
            assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN
