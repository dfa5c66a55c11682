"""Multi-GPU host logic on CPU: the LPT partition and, with world_size 2 over gloo, the
post-forward gather.  (The forward itself has no collective; each rank's shard is validated
against the oracle on that sub-batch in the GPU tests.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fs2_b200
from fs2_b200 import partition


def test_lpt_partition_covers_and_balances(syn):
    lens = syn.random_lengths(512, seed=100)
    for n in (1, 2, 4, 8):
        parts = partition.lpt_partition(lens, n)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(512))
        loads = [sum(partition.utterance_cost(lens[i]) for i in p) for p in parts]
        assert max(loads) / (sum(loads) / n) < 1.01, "LPT should balance 512 utterances to within 1%"
    assert partition.lpt_partition(lens, 3) == partition.lpt_partition(list(lens), 3), "deterministic"


def test_take_repads_to_the_shards_own_max(syn):
    batch = syn.make_batch([30, 7, 19, 12], seed=3)
    sub = partition.take(batch, [1, 3])
    assert sub["max_src_len"] == 12 and tuple(sub["texts"].shape) == (2, 12)
    assert torch.equal(sub["texts"][0, :7], batch["texts"][1, :7])
    assert torch.equal(sub["src_lens"], torch.tensor([7, 12]))


def test_rebalance_plan_is_a_permutation_and_balances_frames(syn):
    lens = syn.random_lengths(512, seed=100)
    g = torch.Generator().manual_seed(1)
    mel_lens = [int(n * (4.0 + 5.0 * float(torch.rand(1, generator=g)))) for n in lens]      # 4 .. 9 frames per phoneme
    for world in (2, 8):
        parts = partition.lpt_partition(lens, world)
        new_parts, moves = partition.plan_rebalance(parts, lens, mel_lens, world)
        assert sorted(i for p in new_parts for i in p) == list(range(512))
        for src in range(world):
            assert sorted(i for dst in range(world) for i in moves[src][dst]) == parts[src]
        for dst in range(world):
            assert sorted(i for src in range(world) for i in moves[src][dst]) == new_parts[dst]
        cost = lambda p: float(sum(partition.stage2_cost(mel_lens[i]) for i in p))
        before = max(cost(p) for p in parts) / (sum(cost(p) for p in parts) / world)
        after = max(cost(p) for p in new_parts) / (sum(cost(p) for p in new_parts) / world)
        assert after < 1.002 and after < before, (before, after)
    assert partition.lpt_by_cost([3, 1, 2, 3], 2) == [[0, 2], [1, 3]]        # ties by index, deterministic


def _rebalance_worker(rank, world, port, lens, mel_lens):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = partition.lpt_partition(lens, world)
        mine = parts[rank]
        L = max(lens[i] for i in mine)
        hidden = torch.zeros(len(mine), L, 256)
        reps = torch.zeros(len(mine), L, dtype=torch.int32)
        for b, gid in enumerate(mine):            # row (utterance g, position j) carries g * 1000 + j + column / 1000
            n = lens[gid]
            hidden[b, :n] = (gid * 1000 + torch.arange(n).float()).unsqueeze(1) + torch.arange(256).float() / 1000
            reps[b, :n] = (gid + torch.arange(n)).int()
        new_parts, moves = partition.plan_rebalance(parts, lens, mel_lens, world)
        h2, r2, l2, ids = partition.exchange_rows(hidden, reps, mine, lens, moves, rank)
        assert sorted(ids) == new_parts[rank]
        assert l2.tolist() == [lens[g] for g in ids] and h2.shape[1] == max(lens[g] for g in ids)
        for b, gid in enumerate(ids):
            n = lens[gid]
            want = (gid * 1000 + torch.arange(n).float()).unsqueeze(1) + torch.arange(256).float() / 1000
            assert torch.equal(h2[b, :n], want) and torch.all(h2[b, n:] == 0)
            assert torch.equal(r2[b, :n], (gid + torch.arange(n)).int()) and torch.all(r2[b, n:] == 0)
    finally:
        dist.destroy_process_group()


def test_row_exchange_over_gloo_world_size_2(syn):
    """The re-partition by frames (partition.plan_rebalance + exchange_rows): every utterance's rows and repeat counts
    arrive bit-exact at their stage-2 owner."""
    lens = syn.random_lengths(21, lo=5, hi=40, seed=6)
    g = torch.Generator().manual_seed(2)
    mel_lens = [int(n * (3.0 + 6.0 * float(torch.rand(1, generator=g)))) for n in lens]
    mp.spawn(_rebalance_worker, args=(2, _free_port(), lens, mel_lens), nprocs=2, join=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lens):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = partition.lpt_partition(lens, world)
        mine = parts[rank]
        t_lens = torch.tensor([3 * lens[i] + rank for i in mine], dtype=torch.int64)
        T = int(t_lens.max())
        local = torch.zeros(len(mine), T, 80)
        for j, i in enumerate(mine):
            local[j, : t_lens[j]] = float(i + 1)
        out, out_lens = partition.gather_padded(local, t_lens, mine, len(lens))
        for r, p in enumerate(parts):
            for i in p:
                n = 3 * lens[i] + r
                assert int(out_lens[i]) == n
                assert torch.all(out[i, :n] == float(i + 1)) and torch.all(out[i, n:] == 0)
    finally:
        dist.destroy_process_group()


def test_gather_over_gloo_world_size_2(syn):
    lens = syn.random_lengths(13, lo=5, hi=40, seed=5)
    mp.spawn(_worker, args=(2, _free_port(), lens), nprocs=2, join=True)


def test_rate_prior_learns_the_conditioning_and_balances_frames(syn):
    """RatePrior: an additive speaking-rate model over (speaker, emotion, arousal, valence) fitted on earlier batches predicts
    the frames of unseen utterances and the LPT partition on it balances the true frame-side cost where the phoneme counts
    do not.  Ground truth here is a synthetic rate with the same structure plus noise (the GPU bench measures the real one)."""
    import numpy as np
    from fs2_b200 import partition
    rng = np.random.default_rng(0)
    eff = [rng.normal(0, s, n) for s, n in ((1.2, 10), (0.6, 5), (0.5, 4), (0.5, 5))]

    def frames_of(batch):
        rate = 5.0 + sum(e[np.asarray(batch[k])] for e, k in zip(eff, ("speakers", "emotions", "arousals", "valences")))
        rate = np.maximum(rate + rng.normal(0, 0.15, len(rate)), 1.0)
        return np.round(rate * np.asarray(batch["src_lens"])).astype(np.int64)

    prior, twin = partition.RatePrior(10, 5, 4, 5), partition.RatePrior(10, 5, 4, 5)
    fresh = syn.config2_batch(seed=0, batch=512)
    assert np.allclose(prior.predict_frames(fresh), np.asarray(fresh["src_lens"]) * partition.FRAMES_PER_PHONEME)
    for seed in (11, 12):
        b = syn.config2_batch(seed=seed)
        seen = frames_of(b)
        prior.observe(b, seen)
        twin.observe(b, seen)
    true = frames_of(fresh)
    pred = prior.predict_frames(fresh)
    assert np.corrcoef(pred, true)[0, 1] > 0.97
    lens = np.asarray(fresh["src_lens"]).tolist()

    def worst(parts):
        cost = [partition.stage2_cost(true[p]).sum() for p in parts]
        return max(cost) / np.mean(cost)

    by_len = partition.lpt_partition(lens, 8)
    by_prior = partition.lpt_partition_by_prior(fresh, 8, prior)
    assert sorted(i for p in by_prior for i in p) == list(range(512))
    assert worst(by_prior) < 1.015 and worst(by_len) > 1.03
    # a second prior fed the same observations plans identically (what makes it usable without communication)
    assert partition.lpt_partition_by_prior(fresh, 8, twin) == by_prior
    with pytest.raises(ValueError):
        bad = dict(fresh)
        bad["speakers"] = np.full(512, 99)
        prior.predict_frames(bad)
