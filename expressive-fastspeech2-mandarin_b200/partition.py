"""Multi-GPU inference = utterance sharding: utterances are independent, every
GPU holds a full weight replica and runs its own context.  This module is the host logic:
  * `lpt_partition`: a length-balanced LPT partition by phoneme counts -- all that is known before stage 1; a forward
    on such shards needs NO collective;
  * `rebalanced_forward`: the same batch re-balanced by FRAMES once stage 1 has produced the durations (the decoder,
    mel_linear and PostNet are 90 % of the work and follow the frames, not the phonemes): one small all-gather of
    mel_lens, a deterministic LPT on the true stage-2 cost, ONE all-to-all of the phoneme rows that change owner
    (P x 1 KB over NVLink), stage 2 where the utterance landed;
  * `RatePrior` + `lpt_partition_by_prior`: the cheap alternative -- predict each utterance's frames BEFORE stage 1 from a
    speaking-rate prior over its conditioning (speaker, emotion, arousal, valence) learned from earlier traffic, and
    balance on that: no collective, no exchange;
  * `gather_padded`: an optional gather of the padded outputs AFTER the forward.
torch.distributed: NCCL over NVLink on the GPU box, gloo (with point-to-point exchange) in the CPU tests.
"""
import heapq

import numpy as np
import torch

FRAMES_PER_PHONEME = 5.0   # planning value only: T_i is unknown before stage 1


def utterance_cost(n_phonemes, frames_per_phoneme=FRAMES_PER_PHONEME):
    """FLOP proxy of one utterance in MFLOP: dense per-row work on both
    sides plus the quadratic attention terms."""
    L = float(n_phonemes)
    T = frames_per_phoneme * L
    return L * 25.43 + 0.004096 * L * L + T * 43.33 + 0.006144 * T * T


def lpt_partition(src_lens, n_parts, frames_per_phoneme=FRAMES_PER_PHONEME):
    """Longest-processing-time-first greedy partition.  Returns n_parts lists of utterance
    indices; every index appears exactly once; ties are broken by index so that every rank
    computes the identical partition without communicating."""
    lens = [int(x) for x in src_lens]
    order = sorted(range(len(lens)), key=lambda i: (-utterance_cost(lens[i], frames_per_phoneme), i))
    parts = [[] for _ in range(n_parts)]
    load = [0.0] * n_parts
    for i in order:
        p = min(range(n_parts), key=lambda k: (load[k], k))
        parts[p].append(i)
        load[p] += utterance_cost(lens[i], frames_per_phoneme)
    for p in parts:
        p.sort()
    return parts


class RatePrior:
    """Speaking-rate prior: frames per phoneme of an utterance ~ a[speaker] + b[emotion] + c[arousal] + d[valence] + e.
    The duration predictor reads the conditioning vector on every row (model/fastspeech2.py:101-110, modules.py:119), so
    an utterance's rate is mostly a property of WHO speaks HOW: on the synthetic weights the additive model fitted on 128
    earlier utterances predicts the frame count of unseen utterances with correlation 0.98, where the phoneme count alone
    reaches 0.61.  `observe` accumulates the normal equations (ridge least squares) from finished batches -- a server
    feeds it its own traffic; `predict_frames` is what `lpt_partition_by_prior` balances on.  Deterministic: ranks that
    observed the same batches hold the same prior and compute the same partition without communicating."""

    def __init__(self, n_speaker, n_emotion, n_arousal, n_valence, ridge=1e-2):
        self.sizes = (int(n_speaker), int(n_emotion), int(n_arousal), int(n_valence))
        self.dim = sum(self.sizes) + 1
        self.xtx = np.zeros((self.dim, self.dim))
        self.xty = np.zeros(self.dim)
        self.ridge = float(ridge)
        self.n_seen = 0
        self._w = None

    def _features(self, batch):
        cols = [np.asarray(batch[k], dtype=np.int64).reshape(-1) for k in ("speakers", "emotions", "arousals", "valences")]
        X = np.zeros((len(cols[0]), self.dim))
        off = 0
        for c, n in zip(cols, self.sizes):
            if (c < 0).any() or (c >= n).any():
                raise ValueError("conditioning index out of range for the rate prior")
            X[np.arange(len(c)), off + c] = 1.0
            off += n
        X[:, -1] = 1.0
        return X

    def observe(self, batch, mel_lens):
        """batch: dict with speakers / emotions / arousals / valences / src_lens (host arrays or CPU tensors);
        mel_lens: the frame counts the forward produced for it."""
        lens = np.asarray(batch["src_lens"], dtype=np.float64).reshape(-1)
        keep = lens > 0
        X = self._features(batch)[keep]
        y = np.asarray(mel_lens, dtype=np.float64).reshape(-1)[keep] / lens[keep]
        self.xtx += X.T @ X
        self.xty += X.T @ y
        self.n_seen += int(keep.sum())
        self._w = None

    def predict_frames(self, batch):
        lens = np.asarray(batch["src_lens"], dtype=np.float64).reshape(-1)
        if self.n_seen == 0:
            return lens * FRAMES_PER_PHONEME
        if self._w is None:
            self._w = np.linalg.solve(self.xtx + self.ridge * np.eye(self.dim), self.xty)
        return np.maximum(self._features(batch) @ self._w, 0.0) * lens


def lpt_partition_by_prior(batch, n_parts, prior):
    """LPT over cost(phonemes) + cost(predicted frames); the same lists on every rank that holds the same prior."""
    lens = np.asarray(batch["src_lens"], dtype=np.float64).reshape(-1)
    cost = lens * 25.43 + 0.004096 * lens * lens + stage2_cost(prior.predict_frames(batch))
    return lpt_by_cost(cost, n_parts)


def stage2_cost(n_frames):
    """FLOP proxy (MFLOP) of the frame side of one utterance: decoder + mel_linear + PostNet
    rows and the quadratic decoder attention."""
    T = np.asarray(n_frames, dtype=np.float64)
    return T * 43.33 + 0.006144 * T * T


def lpt_by_cost(costs, n_parts):
    """Longest-processing-time-first on given costs (ties by index): n_parts sorted index lists, identical on every
    rank.  A heap keeps the greedy step O(log n_parts).  With many items per part (>= 16) the greedy loop is replaced by
    its vectorised cousin -- sort by cost and deal the items out in serpentine order -- which balances a few hundred
    utterances to within a fraction of a percent and costs microseconds instead of a Python loop on the critical path
    between the two stages."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    if n_parts > 1 and len(costs) >= 16 * n_parts:
        k = np.arange(len(order)) % (2 * n_parts)
        dest = np.where(k < n_parts, k, 2 * n_parts - 1 - k)
        return [np.sort(order[dest == p]).tolist() for p in range(n_parts)]
    heap = [(0.0, k) for k in range(n_parts)]
    parts = [[] for _ in range(n_parts)]
    for i in order.tolist():
        load, k = heapq.heappop(heap)
        parts[k].append(i)
        heapq.heappush(heap, (load + float(costs[i]), k))
    for p in parts:
        p.sort()
    return parts


def _exchange(send, send_splits, recv_splits, width, dtype, device, group):
    """All-to-all of row blocks: `send` [sum(send_splits), width] ordered by destination rank.  NCCL: one
    all_to_all_single; gloo (CPU tests) has no all-to-all: point-to-point isend / irecv with the same semantics."""
    import torch.distributed as dist
    recv = torch.empty(int(sum(recv_splits)), width, dtype=dtype, device=device)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, list(recv_splits), list(send_splits), group=group)
        return recv
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    so, ro = np.concatenate(([0], np.cumsum(send_splits))), np.concatenate(([0], np.cumsum(recv_splits)))
    recv[ro[rank]: ro[rank + 1]] = send[so[rank]: so[rank + 1]]
    ops = []
    for r in range(world):
        if r == rank:
            continue
        if send_splits[r]:
            ops.append(dist.P2POp(dist.isend, send[so[r]: so[r + 1]].contiguous(), r, group))
        if recv_splits[r]:
            ops.append(dist.P2POp(dist.irecv, recv[ro[r]: ro[r + 1]], r, group))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return recv


class Moves:
    """Who sends what to whom: owner[i] = stage-1 rank of utterance i, dest[i] = its stage-2 rank.  moves[src][dst] is the
    sorted list of utterance ids going from src to dst (computed on demand: a rank only needs its own row and column)."""

    def __init__(self, owner, dest, n_parts):
        self.owner, self.dest, self.n = np.asarray(owner), np.asarray(dest), n_parts

    def __len__(self):
        return self.n

    def __getitem__(self, src):
        sel = self.owner == src
        return [np.nonzero(sel & (self.dest == d))[0].tolist() for d in range(self.n)]

    def column(self, dst):
        sel = self.dest == dst
        return [np.nonzero(sel & (self.owner == s))[0].tolist() for s in range(self.n)]


def plan_rebalance(parts, src_lens, mel_lens, n_parts):
    """Given the phoneme shards `parts` (owner of every utterance during stage 1), all src_lens and the mel_lens stage 1
    produced, returns (new_parts, moves): new_parts = LPT of the TRUE stage-2 cost; moves[src][dst] = sorted utterance
    ids that rank src sends to rank dst (src == dst: stay).  Pure host arithmetic, identical on every rank."""
    new_parts = lpt_by_cost(stage2_cost(mel_lens), n_parts)
    owner = np.empty(len(src_lens), dtype=np.int64)
    dest = np.empty(len(src_lens), dtype=np.int64)
    for r, p in enumerate(parts):
        owner[p] = r
    for r, p in enumerate(new_parts):
        dest[p] = r
    return new_parts, Moves(owner, dest, n_parts)


def exchange_rows(hidden, reps, local_ids, src_lens_all, moves, rank, group=None):
    """Move the exported stage-1 rows to their stage-2 owners.  hidden [B_local, L_local, 256] fp32 and reps
    [B_local, L_local] int32 are this rank's stage-1 export for the utterances `local_ids` (sorted global ids);
    src_lens_all: every utterance's phoneme count (host list).  Returns (hidden', reps', src_lens', ids') for the
    utterances this rank now owns, padded to their own maximum, ids' in the order of the rows."""
    world = len(moves)
    dev = hidden.device
    lens = np.asarray(src_lens_all, dtype=np.int64)
    pos = {g: j for j, g in enumerate(local_ids)}
    # send side: for every destination the (local utterance, position) pairs of the real rows, in id order
    send_ids = moves[rank]
    flat = [g for ids in send_ids for g in ids]
    if flat:
        fl = lens[flat]
        b_idx = np.repeat(np.array([pos[g] for g in flat], dtype=np.int64), fl)
        j_idx = np.arange(int(fl.sum())) - np.repeat(np.cumsum(fl) - fl, fl)       # position inside each utterance
    else:
        b_idx = j_idx = np.zeros(0, dtype=np.int64)
    bi, ji = torch.from_numpy(b_idx).to(dev, non_blocking=True), torch.from_numpy(j_idx).to(dev, non_blocking=True)
    # one message per peer: 256 fp32 columns of the row + its repeat count, bit-cast into a 257th column
    send = torch.cat([hidden[bi, ji], reps[bi, ji].contiguous().view(torch.float32).unsqueeze(1)], dim=1)
    send_splits = [int(lens[ids].sum()) if ids else 0 for ids in send_ids]
    recv_ids = moves.column(rank) if isinstance(moves, Moves) else [moves[s][rank] for s in range(world)]
    recv_splits = [int(lens[ids].sum()) if ids else 0 for ids in recv_ids]
    recv = send if world == 1 else _exchange(send, send_splits, recv_splits, 257, torch.float32, dev, group)
    ids = [g for r_ids in recv_ids for g in r_ids]
    my_lens = lens[ids] if ids else np.zeros(0, dtype=np.int64)
    B, L = len(ids), int(my_lens.max()) if len(ids) else 0
    out_h = torch.zeros(B, L, 256, dtype=torch.float32, device=dev)
    out_r = torch.zeros(B, L, dtype=torch.int32, device=dev)
    if B:
        rb = torch.from_numpy(np.repeat(np.arange(B), my_lens)).to(dev, non_blocking=True)
        rj = torch.from_numpy(np.arange(int(my_lens.sum())) - np.repeat(np.cumsum(my_lens) - my_lens, my_lens)).to(dev, non_blocking=True)
        out_h[rb, rj] = recv[:, :256]
        out_r[rb, rj] = recv[:, 256].contiguous().view(torch.int32)
    return out_h, out_r, torch.from_numpy(my_lens).to(dev), ids


def rebalanced_forward(model, batch, parts, rank, group=None, p_control=1.0, e_control=1.0, d_control=1.0):
    """One sharded forward with the frames re-balanced after stage 1.  `batch` is the WHOLE batch (host tensors, known to
    every rank), `parts` its phoneme partition (`lpt_partition`).  Returns a dict: `ids` (global ids of the utterances this
    rank decoded), `mel`, `postnet`, `mel_mask`, `mel_lens` for them, `stage1` (this rank's phoneme-side outputs for
    parts[rank]) and `new_parts`.  Collectives: one all_gather of mel_lens, one all-to-all of rows."""
    import torch.distributed as dist
    world = len(parts)
    dev = model._device()
    mine = parts[rank]
    sub = take(batch, mine)
    s1 = model.encode(*[sub[k].to(dev, non_blocking=True) for k in ("speakers", "emotions", "arousals", "valences", "texts",
                                                                       "src_lens")], sub["max_src_len"],
                      p_control=p_control, e_control=e_control, d_control=d_control)
    n = len(batch["src_lens"])
    b_max = max(len(p) for p in parts)
    local = torch.full((b_max,), -1, dtype=torch.int64, device=dev)
    local[: len(mine)] = s1["mel_lens"]
    if world > 1:
        gathered = torch.empty(world * b_max, dtype=torch.int64, device=dev)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(gathered, local, group=group)
        else:
            chunks = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(chunks, local, group=group)
            gathered = torch.cat(chunks)
    else:
        gathered = local
    g = gathered.cpu().numpy().reshape(world, b_max)          # (the only host sync besides the two stage read-backs)
    mel_lens = np.zeros(n, dtype=np.int64)
    for r, p in enumerate(parts):
        mel_lens[p] = g[r, : len(p)]
    new_parts, moves = plan_rebalance(parts, batch["src_lens"].tolist(), mel_lens, world)
    hidden, reps, lens, ids = exchange_rows(s1["hidden"], s1["reps"], mine, batch["src_lens"].tolist(), moves, rank, group)
    out = {"ids": ids, "stage1": s1, "new_parts": new_parts, "mel_lens_all": mel_lens}
    if ids:
        out["mel"], out["postnet"], out["mel_mask"], out["mel_lens"] = model.decode(hidden, reps, lens)
    return out


def take(batch, indices):
    """The sub-batch of `indices`, re-padded to its own max_src_len (each shard is validated
    against the reference run on that sub-batch: PostNet / predictor tails depend on the
    batch's own L_max and T_max -- SURVEY.md §8e)."""
    idx = torch.as_tensor(indices, dtype=torch.int64)
    lens = batch["src_lens"].index_select(0, idx)
    L = int(lens.max()) if len(indices) else 0
    out = {k: batch[k].index_select(0, idx) for k in ("speakers", "emotions", "arousals", "valences")}
    out["texts"] = batch["texts"].index_select(0, idx)[:, :L].contiguous()
    out["src_lens"] = lens
    out["max_src_len"] = L
    return out


def gather_padded(local, local_lens, local_indices, total, group=None):
    """All-gather per-utterance outputs ([B_local, T_local, C] + lengths) into the original
    utterance order on every rank: one MAX all-reduce of the padded length, one all_gather of
    the sizes and one of the padded tensors.  Runs AFTER the forward; never inside it."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = local.device
    meta = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    b_max = max(int(m[0]) for m in metas)
    t_max = max(int(m[1]) for m in metas)
    C = local.shape[2]
    pad = torch.zeros(b_max, t_max, C, dtype=local.dtype, device=dev)
    pad[: local.shape[0], : local.shape[1]] = local
    lens = torch.zeros(b_max, dtype=torch.int64, device=dev)
    lens[: local.shape[0]] = local_lens
    idx = torch.full((b_max,), -1, dtype=torch.int64, device=dev)
    idx[: local.shape[0]] = torch.as_tensor(local_indices, dtype=torch.int64, device=dev)
    pads = [torch.zeros_like(pad) for _ in range(world)]
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    all_idx = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(pads, pad, group=group)
    dist.all_gather(all_lens, lens, group=group)
    dist.all_gather(all_idx, idx, group=group)
    out = torch.zeros(total, t_max, C, dtype=local.dtype, device=dev)
    out_lens = torch.zeros(total, dtype=torch.int64, device=dev)
    for p, l, i in zip(pads, all_lens, all_idx):
        keep = i >= 0
        out[i[keep]] = p[keep]
        out_lens[i[keep]] = l[keep]
    return out, out_lens
