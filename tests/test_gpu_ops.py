"""Single-operator parity on the GPU, each through its C-ABI entry point, against a plain
torch float64 statement of the same op on the CPU (floating point) or the oracle's integer
helpers (bit-exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import fs2_oracle as O
from gpu_util import DEV, lib, ptr, round_tf32, stream

pytestmark = pytest.mark.gpu


def conv_ref(A, W, bias, pad, act, residual, vpos, room, extra):
    """out[r] = act(sum_t A[r+t-pad] @ W[t].T + bias) (+residual), masked rows -> 0; float64."""
    rows, K = A.shape
    taps, N, _ = W.shape
    A64 = torch.zeros(rows + 2 * taps, K, dtype=torch.float64)
    A64[taps: taps + rows] = A.double()
    out = bias.double().unsqueeze(0).repeat(rows, 1)
    for t in range(taps):
        lo = taps + t - pad
        out += A64[lo: lo + rows] @ W[t].double().T
    if act == 1:
        out = torch.relu(out)
    elif act == 2:
        out = torch.tanh(out)
    if residual is not None:
        out = out + residual.double()
    if vpos is not None:
        live = vpos < torch.minimum(torch.full_like(room, extra), room)
        out = out * live.unsqueeze(1)
    return out


CASES = [
    # rows, K, N, taps, act, residual, mask, name
    (300, 256, 768, 1, 0, False, False, "qkv"),
    (517, 256, 1024, 9, 1, False, False, "ffn_conv9"),
    (260, 1024, 256, 1, 0, True, False, "ffn_w2"),
    (131, 256, 256, 3, 1, False, False, "predictor_conv3"),
    (200, 256, 80, 1, 0, False, True, "mel_linear"),
    (333, 80, 512, 5, 2, False, True, "postnet_first"),
    (333, 512, 512, 5, 2, False, True, "postnet_mid"),
    (129, 512, 80, 5, 0, True, True, "postnet_last"),
    (5, 256, 256, 1, 0, False, False, "tiny"),
    # single row tile + long K loop: the K-split cluster form (8 CTAs per column tile, partial tiles through the workspace)
    (56, 256, 1024, 9, 1, False, False, "one_tile_conv9_ksplit"),
    (100, 512, 512, 5, 2, False, True, "one_tile_postnet_mid_ksplit"),
    (128, 512, 80, 5, 0, True, True, "one_tile_postnet_last_ksplit"),
    (77, 1024, 256, 1, 0, True, False, "one_tile_w2_ksplit"),
    (40000, 256, 768, 1, 0, False, False, "qkv_many_tiles"),
    (30011, 1024, 256, 1, 0, True, True, "w2_many_tiles"),
    (25000, 512, 80, 5, 0, True, True, "postnet_last_many_tiles"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[-1] for c in CASES])
def test_conv_gemm(case):
    rows, K, N, taps, act, use_res, use_mask, _ = case
    g = torch.Generator().manual_seed(rows * 7 + K + N + taps)
    A = round_tf32(torch.randn(rows, K, generator=g))
    W = round_tf32(torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps))
    bias = torch.randn(N, generator=g)
    res = torch.randn(rows, N, generator=g) if use_res else None
    vpos = torch.randint(-3, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    room = torch.randint(0, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    extra = 2
    want = conv_ref(A, W, bias, (taps - 1) // 2, act, res, vpos, room, extra)

    dA, dW, db = A.to(DEV), W.to(DEV), bias.to(DEV)
    dres = res.to(DEV) if use_res else None
    dv = vpos.to(DEV) if use_mask else None
    dr = room.to(DEV) if use_mask else None
    out = torch.full((rows, N), float("nan"), device=DEV)
    code = lib().fs2_op_conv_gemm(stream(), 0, ptr(dA), K, rows, ptr(dW), ptr(db), taps, (taps - 1) // 2, K, N,
                                  act, ptr(dres), N, ptr(dv), ptr(dr), extra, ptr(out), N)
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    # operands are TF32-exact, so only the fp32 accumulation order differs
    assert err < 2e-4, f"max abs err {err}"


LN_CASES = [
    # rows, K, taps, act, residual, store C, head, name
    (300, 256, 1, 0, True, True, False, "fc_residual_ln"),
    (1000, 1024, 1, 0, True, True, False, "w2_residual_ln"),
    (137, 256, 3, 1, False, True, False, "predictor_conv1_relu_ln"),
    (137, 256, 3, 1, False, False, True, "predictor_conv2_relu_ln_head"),
    (20000, 256, 1, 0, True, True, True, "many_tiles"),
]


@pytest.mark.parametrize("case", LN_CASES, ids=[c[-1] for c in LN_CASES])
def test_conv_gemm_fused_layernorm(case):
    """LayerNorm(act(conv + bias) + residual) fused into the GEMM epilogue."""
    rows, K, taps, act, use_res, store, use_head, _ = case
    g = torch.Generator().manual_seed(rows + K + taps)
    A = round_tf32(torch.randn(rows, K, generator=g))
    W = round_tf32(torch.randn(taps, 256, K, generator=g) / np.sqrt(K * taps))
    bias = torch.randn(256, generator=g) * 0.3
    res = torch.randn(rows, 256, generator=g) if use_res else None
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2
    hw, hb = torch.randn(256, generator=g) / 16, torch.randn(1, generator=g)
    vpos = torch.randint(-3, 3, (rows,), generator=g, dtype=torch.int32)
    room = torch.randint(0, 3, (rows,), generator=g, dtype=torch.int32)
    extra = 1
    pre = conv_ref(A, W, bias, (taps - 1) // 2, act, res, None, None, 0)
    y = torch.nn.functional.layer_norm(pre, (256,), gamma.double(), beta.double(), 1e-5)
    live = vpos < torch.minimum(torch.full_like(room, extra), room)
    dot = y @ hw.double() + hb.double()
    y = y * live.unsqueeze(1)
    d = lambda t: t.to(DEV) if t is not None else None
    dA, dW, db, dres, dg, dbe, dv, dr, dhw, dhb = map(d, (A, W, bias, res, gamma, beta, vpos, room, hw, hb))
    out = torch.full((rows, 256), float("nan"), device=DEV) if store else None
    head = torch.zeros(rows, device=DEV) if use_head else None
    code = lib().fs2_op_conv_gemm_ln(stream(), ptr(dA), K, rows, ptr(dW), ptr(db), taps, (taps - 1) // 2, K, act,
                                     ptr(dres), 256, ptr(dg), ptr(dbe), ptr(dv), ptr(dr), extra, ptr(out), 256,
                                     ptr(dhw) if use_head else None, ptr(dhb) if use_head else None, ptr(head))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    if store:
        got = out.cpu().double()
        assert torch.isfinite(got).all()
        assert (got - y).abs().max().item() < 3e-4
    if use_head:
        got = head.cpu().double()
        assert (got[live] - dot[live]).abs().max().item() < 3e-4
        assert (got[~live] == 0).all()


@pytest.mark.parametrize("pair", [0, 1, 2, 3, 4],
                         ids=["one_cta_per_tile", "paired_kv_multicast", "two_sm_pair", "persistent_one_group", "persistent"])
@pytest.mark.parametrize("lens", [[1], [5, 64, 65, 33], [200, 7, 129, 128, 127], [700], [0, 3, 0, 300, 1],
                                  [37] * 70 + [513, 2, 1024]])
def test_attention(lens, pair):
    """The forms of the kernel: one CTA per 128-query tile; the persistent kernel that walks the work list (the default); clusters of two CTAs on adjacent query tiles that share every K/V
    tile through TMA multicast (odd tile counts leave a loads-only CTA); and the 2-SM kernel (attention_tc2.cuh: M = 256
    cta_group::2 MMAs, 128-key tiles split between the CTAs; odd tile counts leave a CTA working on nobody's rows)."""
    lib().fs2_debug_set_flag(8, 0 if pair >= 3 else pair)
    lib().fs2_debug_set_flag(10, {3: 2, 4: 3}.get(pair, 0))   # attention_tcp.cuh / attention_tcq.cuh (the default) at every size
    try:
        _attention_case(lens)
    finally:
        lib().fs2_debug_set_flag(8, -1)
        lib().fs2_debug_set_flag(10, 3)


def _persistent_lens(kind):
    g = torch.Generator().manual_seed(11)
    if kind == "config2_like":
        return [int(x) for x in torch.randint(80, 701, (64,), generator=g)]
    if kind == "many_short_with_empties":
        return [int(x) for x in torch.randint(0, 130, (400,), generator=g)]
    if kind == "one_item_more_than_sms":       # 75 single-tile utterances -> 150 (tile, head) items on 148 CTAs
        return [int(x) for x in torch.randint(1, 129, (75,), generator=g)]
    return [37] * 70 + [513, 2, 1024] + [64, 65, 127, 128, 129, 191, 192, 193] * 3


@pytest.mark.parametrize("form", [2, 3], ids=["one_group", "softmax_and_accumulate_groups"])
@pytest.mark.parametrize("kind", ["config2_like", "many_short_with_empties", "one_item_more_than_sms", "mixed_boundaries"])
def test_attention_persistent(kind, form):
    """attention_tcp.cuh / attention_tcq.cuh: one CTA per SM walks the work list with Q travelling through the K ring and the
    pipelines running across items (tcq: the row work split between a softmax and an accumulate warpgroup).  Same products
    in the same order as the one-CTA-per-item kernel: the outputs are bit-identical to it (debug flag 10 = 0) and within the
    TF32 tolerance of float64."""
    lens = _persistent_lens(kind)
    try:
        lib().fs2_debug_set_flag(10, form)
        got = _attention_case(lens)
        lib().fs2_debug_set_flag(10, 0)
        ref = _attention_case(lens)
    finally:
        lib().fs2_debug_set_flag(10, 3)
    assert torch.equal(got, ref)


def _attention_case(lens):
    g = torch.Generator().manual_seed(sum(lens))
    gap = 4
    starts, r = [], gap
    for n in lens:
        starts.append(r)
        r += n + gap
    rows = r
    qkv = torch.randn(rows, 768, generator=g)
    want = torch.zeros(rows, 256, dtype=torch.float64)
    for s, n in zip(starts, lens):
        x = qkv[s: s + n].double()
        for h in range(2):
            q, k, v = (x[:, i * 256 + h * 128: i * 256 + (h + 1) * 128] for i in range(3))
            p = torch.softmax(q @ k.T / np.sqrt(128.0), dim=1)
            want[s: s + n, h * 128: (h + 1) * 128] = p @ v
    dq = qkv.to(DEV)
    out = torch.zeros(rows, 256, device=DEV)
    ds = torch.tensor(starts, dtype=torch.int32, device=DEV)
    dl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    code = lib().fs2_op_attention(stream(), ptr(dq), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    live = torch.zeros(rows, dtype=torch.bool)
    for s, n in zip(starts, lens):
        live[s: s + n] = True
    assert (got[~live] == 0).all(), "rows outside the utterances must not be written"
    err = (got[live] - want[live]).abs().max().item()
    assert err < 5e-3, f"max abs err {err}"   # TF32 operands on unit-variance data, d_k = 128
    return got


X3_CASES = [c for c in CASES if c[-1] in ("qkv", "ffn_conv9", "ffn_w2", "postnet_first", "postnet_last", "w2_many_tiles")]


@pytest.mark.parametrize("case", X3_CASES, ids=[c[-1] for c in X3_CASES])
def test_conv_gemm_split_operands(case):
    """FS2_MATH_TF32X3 (the `parity` mode): unrounded fp32 operands, three TF32 terms -> fp32-class accuracy."""
    rows, K, N, taps, act, use_res, use_mask, _ = case
    g = torch.Generator().manual_seed(rows * 5 + K + N + taps)
    A = torch.randn(rows, K, generator=g)
    W = torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps)
    bias = torch.randn(N, generator=g)
    res = torch.randn(rows, N, generator=g) if use_res else None
    vpos = torch.randint(-3, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    room = torch.randint(0, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    want = conv_ref(A, W, bias, (taps - 1) // 2, act, res, vpos, room, 2)
    d = lambda t: t.to(DEV) if t is not None else None
    dA, dW, db, dres, dv, dr = map(d, (A, W, bias, res, vpos, room))
    out = torch.full((rows, N), float("nan"), device=DEV)
    code = lib().fs2_op_conv_gemm(stream(), 2, ptr(dA), K, rows, ptr(dW), ptr(db), taps, (taps - 1) // 2, K, N,
                                  act, ptr(dres), N, ptr(dv), ptr(dr), 2, ptr(out), N)
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    # the same contraction in plain TF32 (operands rounded once): the split form must be far closer to float64
    plain = torch.full((rows, N), float("nan"), device=DEV)
    code = lib().fs2_op_conv_gemm(stream(), 0, ptr(round_tf32(A).to(DEV)), K, rows, ptr(round_tf32(W).to(DEV)), ptr(db), taps,
                                  (taps - 1) // 2, K, N, act, ptr(dres), N, ptr(dv), ptr(dr), 2, ptr(plain), N)
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    err_tf32 = (plain.cpu().double() - want).abs().max().item()
    print(f"split-operand max abs err {err:.2e}, plain TF32 {err_tf32:.2e}")
    # what remains is the tensor core's accumulation (the dropped a_lo*w_lo term is ~2^-22 relative): measured <= 1.4e-4
    # on the K = 2304 conv, against ~2e-3 for plain TF32 on the same operands
    assert err < 3e-4, f"max abs err {err}"
    assert err < err_tf32 / 4, (err, err_tf32)


@pytest.mark.parametrize("d_control", [1.0, 1.5, 0.5, 2.0])
def test_durations_bit_exact(d_control):
    g = torch.Generator().manual_seed(11)
    B, L = 9, 70
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    log_d = torch.randn(B, L, generator=g) * 0.8 + 1.2
    # exact .5 ties and negative values
    log_d[1, :4] = torch.log(torch.tensor([1.5, 2.5, 3.5, 0.4]))
    pad = O.pad_mask(lens, L)
    log_d = log_d.masked_fill(pad, 0.0)
    want_d = O.duration_rounded(log_d, d_control)
    want_reps = O.repeat_counts(want_d) * (~pad)
    want_cum = torch.cumsum(want_reps, 1).to(torch.int32)
    d_round = torch.zeros(B, L, device=DEV)
    cum = torch.zeros(B, L, dtype=torch.int32, device=DEV)
    mel_lens = torch.zeros(B, dtype=torch.int64, device=DEV)
    dl, dlens = log_d.to(DEV), lens.to(DEV)
    code = lib().fs2_op_durations(stream(), ptr(dl), 0, d_control, ptr(dlens), B, L, ptr(d_round), ptr(cum), ptr(mel_lens))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    # expf on the GPU and torch.exp on the CPU may differ in the last ulp; a different integer is
    # only tolerated (and reported) at a rounding boundary
    got_d = d_round.cpu()
    diff = (got_d != want_d)
    if diff.any():
        frac = (torch.exp(log_d.double()) - 1) % 1.0
        assert ((frac[diff] - 0.5).abs() < 1e-5).all(), "duration mismatch away from a .5 rounding boundary"
        print("rounding-boundary ties:", int(diff.sum()))
    else:
        assert torch.equal(cum.cpu(), want_cum)
        assert torch.equal(mel_lens.cpu(), want_reps.sum(1))

    # target durations (ints and fractions) expand by truncation
    tgt = torch.randint(0, 12, (B, L), generator=g).float() * (~pad)
    tgt[2, 0] = 2.75
    dt = tgt.to(DEV)
    code = lib().fs2_op_durations(stream(), ptr(dt), 1, 1.0, ptr(dlens), B, L, None, ptr(cum), ptr(mel_lens))
    assert code == 0
    torch.cuda.synchronize()
    reps = O.repeat_counts(tgt)
    assert torch.equal(cum.cpu(), torch.cumsum(reps, 1).to(torch.int32))
    assert torch.equal(mel_lens.cpu(), reps.sum(1))


def test_bucketize_bit_exact():
    g = torch.Generator().manual_seed(5)
    bins = torch.linspace(-2.5, 9.0, 255)
    v = torch.randn(40000, generator=g) * 3 + 2
    v[:255] = bins                       # values equal to a boundary take that boundary's index
    v[255:510] = torch.nextafter(bins, torch.tensor(float("inf")))
    v[510] = float("nan")
    v[511] = float("inf")
    v[512] = -float("inf")
    want = O.bucket_index(v, bins).to(torch.int32)
    dv, db = v.to(DEV), bins.to(DEV)
    idx = torch.zeros(v.numel(), dtype=torch.int32, device=DEV)
    code = lib().fs2_op_bucketize(stream(), ptr(dv), v.numel(), ptr(db), 255, ptr(idx))
    assert code == 0
    torch.cuda.synchronize()
    assert torch.equal(idx.cpu(), want)


def test_frame_map_bit_exact():
    g = torch.Generator().manual_seed(8)
    B, L = 7, 50
    reps = torch.randint(0, 9, (B, L), generator=g)
    reps[3] = 0
    reps[4, 10:] = 0
    cum = torch.cumsum(reps, 1).to(torch.int32)
    T = int(cum[:, -1].max())
    want = torch.full((B, T), -1, dtype=torch.int32)
    for b in range(B):
        m = O.frame_to_phoneme_map(reps[b].numpy())
        want[b, : len(m)] = torch.from_numpy(m).to(torch.int32)
    dc = cum.to(DEV)
    out = torch.zeros(B, T, dtype=torch.int32, device=DEV)
    code = lib().fs2_op_frame_map(stream(), ptr(dc), B, L, T, ptr(out))
    assert code == 0
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("rows", [5, 128, 300, 1500, 4097, 9000, 19500, 26788, 37900])
def test_ffn_fused(rows):
    """conv9 -> ReLU -> w2 -> +x -> LayerNorm -> mask in one kernel (hidden rows stay in tensor memory) against float64
    with the hidden activations rounded to TF32 exactly as the kernel rounds them (nearest, ties away)."""
    g = torch.Generator().manual_seed(rows)
    x = round_tf32(torch.randn(rows, 256, generator=g))
    w1 = round_tf32(torch.randn(9, 1024, 256, generator=g) / np.sqrt(256 * 9))
    b1 = torch.randn(1024, generator=g) * 0.3
    w2 = round_tf32(torch.randn(1, 256, 1024, generator=g) / np.sqrt(1024))
    b2 = torch.randn(256, generator=g) * 0.3
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2
    vpos = torch.randint(-4, 2, (rows,), generator=g, dtype=torch.int32)
    room = torch.randint(0, 3, (rows,), generator=g, dtype=torch.int32)
    hid = conv_ref(x, w1, b1, 4, 1, None, None, None, 0)
    hid = round_tf32(hid.float()).double()          # the kernel sees fp32 accumulations: round those
    pre = hid @ w2[0].double().T + b2.double() + x.double()
    want = torch.nn.functional.layer_norm(pre, (256,), gamma.double(), beta.double(), 1e-5)
    live = vpos < torch.minimum(torch.zeros_like(room), room)
    want = want * live.unsqueeze(1)
    d = lambda t: t.to(DEV)
    dx, dw1, db1, dw2, db2, dg, dbe, dv, dr = map(d, (x, w1, b1, w2, b2, gamma, beta, vpos, room))
    out = torch.full((rows, 256), float("nan"), device=DEV)
    code = lib().fs2_op_ffn_fused(stream(), ptr(dx), rows, ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(dg), ptr(dbe), ptr(dv),
                                  ptr(dr), 0, ptr(out))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    # hidden values that sit on a TF32 rounding boundary may round the other way after a different summation order:
    # one such flip moves an output by ~2^-11 * |w2| ~ 1e-5
    assert err < 1e-3, f"max abs err {err}"


@pytest.mark.parametrize("lens", [[300], [130, 257]])
def test_attention_growing_scores(lens):
    """Scores that keep growing along the key axis: the running maximum of the online softmax moves at every key tile,
    by tens of units of the raw score (a lazily rescaled accumulator in TMEM was tried on this test and showed no speed-up)."""
    g = torch.Generator().manual_seed(5)
    gap = 4
    starts, r = [], gap
    for n in lens:
        starts.append(r)
        r += n + gap
    rows = r
    qkv = torch.randn(rows, 768, generator=g)
    for s, n in zip(starts, lens):
        ramp = (1.0 + torch.arange(n).float() / 16.0).unsqueeze(1)       # key t scaled by 1 + t/16: maxima jump tile to tile
        qkv[s: s + n, 256:512] *= ramp
        qkv[s: s + n, 0:256] = qkv[s: s + n, 0:256].abs() * 0.5 + 0.5      # positive queries ...
        qkv[s: s + n, 256:512] = qkv[s: s + n, 256:512].abs()              # ... and keys: scores grow monotonically
    want = torch.zeros(rows, 256, dtype=torch.float64)
    for s, n in zip(starts, lens):
        x = qkv[s: s + n].double()
        for h in range(2):
            q, k, v = (x[:, i * 256 + h * 128: i * 256 + (h + 1) * 128] for i in range(3))
            sc = q @ k.T / np.sqrt(128.0)
            assert float(sc.max(dim=1).values.min()) > 60, "the row maxima must be far above the first tile's"
            want[s: s + n, h * 128: (h + 1) * 128] = torch.softmax(sc, dim=1) @ v
    dq = qkv.to(DEV)
    out = torch.zeros(rows, 256, device=DEV)
    ds = torch.tensor(starts, dtype=torch.int32, device=DEV)
    dl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    code = lib().fs2_op_attention(stream(), ptr(dq), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    live = torch.zeros(rows, dtype=torch.bool)
    for s, n in zip(starts, lens):
        live[s: s + n] = True
    assert torch.isfinite(got).all()
    # sharply peaked softmax over TF32-rounded scores of magnitude ~1e3: the weights move by up to 2^-11 * |score| relative
    err = (got[live] - want[live]).abs().max().item()
    assert err < 0.5, f"max abs err {err}"
    rel = ((got[live] - want[live]).abs().sum() / want[live].abs().sum()).item()
    assert rel < 0.1, f"relative L1 error {rel}"


@pytest.mark.parametrize("shape", [(25765, 256, 1024, 9, 1), (12900, 512, 512, 5, 2), (257, 256, 1024, 9, 1)],
                         ids=["conv9_odd_tiles", "postnet_mid", "two_tiles_plus_one_row"])
def test_conv_gemm_two_sm(shape):
    """cta_group::2 MMAs (one 256-row MMA per CTA pair, each CTA holding half of the weight tile) against float64 and
    against the single-CTA MMA form (debug flag 6 = 0): same products, same accumulation order per row -> bit-equal."""
    rows, K, N, taps, act = shape
    g = torch.Generator().manual_seed(rows + K)
    A = round_tf32(torch.randn(rows, K, generator=g))
    W = round_tf32(torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps))
    bias = torch.randn(N, generator=g)
    dA, dW, db = A.to(DEV), W.to(DEV), bias.to(DEV)
    outs = []
    L = lib()
    try:
        for flag in (1, 0):
            L.fs2_debug_set_flag(6, flag)
            out = torch.full((rows, N), float("nan"), device=DEV)
            code = L.fs2_op_conv_gemm(stream(), 0, ptr(dA), K, rows, ptr(dW), ptr(db), taps, (taps - 1) // 2, K, N, act,
                                      None, N, None, None, 0, ptr(out), N)
            assert code == 0, L.fs2_last_error(None)
            torch.cuda.synchronize()
            outs.append(out.cpu())
    finally:
        L.fs2_debug_set_flag(6, 1)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]), "2-SM and 1-SM forms must agree bit for bit"
    idx = torch.cat([torch.arange(0, min(rows, 300)), torch.arange(max(rows - 300, 0), rows)]).unique()
    want = conv_ref(A, W, bias, (taps - 1) // 2, act, None, None, None, 0)[idx]
    assert (outs[0][idx].double() - want).abs().max().item() < 2e-4
