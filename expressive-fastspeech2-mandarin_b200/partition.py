"""Multi-GPU inference = utterance sharding: utterances are independent, every
GPU holds a full weight replica and runs its own context, and there is NO collective on the hot
path.  This module is the host logic: a length-balanced LPT partition and an optional gather of
the padded outputs over torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
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
