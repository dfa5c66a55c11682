"""FS2_MATH_BF16 mode of the CUDA path (north star: "a separately stated tolerance for the bf16 mode").

Every Conv1d/Linear runs on tcgen05 kind::f16 with bf16 operands (weights rounded once, activations written as
bf16 by the producing epilogue); accumulation, the residual stream, the attention (TF32), LayerNorm, softmax and
all outputs stay fp32.

Stated tolerances, max-abs on valid rows against the fp64 reference (SURVEY.md §8(c) budget for bf16 operands:
predictions 1.3e-2, mel 8.7e-3, postnet 1.0e-2, mean 1.6e-3):
    log-duration / pitch / energy <= 1.5e-2;  mel / postnet mel <= 1.2e-2, mean-abs <= 2.2e-3
(measured on B200: predictions <= 1.02e-2, mel / postnet <= 9.3e-3, mean 1.6e-3 -- profiles/r02_parity_*.txt).
Integer stages are the same kernels as in TF32 mode and stay bit-exact given equal inputs (teacher forcing).
Operator level: against float64 on the SAME bf16-rounded operands only the fp32 accumulation order differs.
"""
import numpy as np
import pytest
import torch

from gpu_util import DEV, err_stats, lib, model_for, ptr, run, stream
from helpers import load_golden, valid_rows
from test_gpu_forward import log_diag, teacher_kwargs
from test_gpu_ops import conv_ref

pytestmark = pytest.mark.gpu

TOL_PRED = 1.5e-2
TOL_MEL_MAX = 1.2e-2
TOL_MEL_MEAN = 2.2e-3

CASES = [
    # rows, K, N, taps, act, residual, mask, ln, fp32 out, bf16 out, name
    (300, 256, 768, 1, 0, False, False, False, True, False, "qkv_f32_out"),
    (517, 256, 1024, 9, 1, False, False, False, False, True, "ffn_conv9_bf16_out"),
    (260, 1024, 256, 1, 0, True, True, True, True, True, "ffn_w2_ln_dual_out"),
    (131, 256, 256, 3, 1, False, True, True, False, True, "predictor_conv1_ln_bf16_out"),
    (200, 256, 80, 1, 0, False, True, False, True, True, "mel_linear_dual_out"),
    (333, 80, 512, 5, 2, False, True, False, False, True, "postnet_first"),
    (333, 512, 512, 5, 2, False, True, False, False, True, "postnet_mid"),
    (129, 512, 80, 5, 0, True, True, False, True, False, "postnet_last"),
    (5, 256, 256, 1, 0, True, False, True, True, True, "tiny_ln"),
    (30011, 1024, 256, 1, 0, True, True, True, True, True, "w2_ln_many_tiles"),
    (26000, 256, 1024, 9, 1, False, False, False, False, True, "conv9_many_tiles"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[-1] for c in CASES])
def test_conv_gemm_bf16(case):
    rows, K, N, taps, act, use_res, use_mask, ln, out32, out16, _ = case
    g = torch.Generator().manual_seed(rows * 5 + K + N + taps)
    A = torch.randn(rows, K, generator=g).bfloat16()
    W = (torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps)).bfloat16()
    bias = torch.randn(N, generator=g) * 0.3
    res = torch.randn(rows, N, generator=g) if use_res else None
    gamma = torch.rand(N, generator=g) + 0.5 if ln else None
    beta = torch.randn(N, generator=g) * 0.2 if ln else None
    vpos = torch.randint(-3, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    room = torch.randint(0, 4, (rows,), generator=g, dtype=torch.int32) if use_mask else None
    extra = 2
    want = conv_ref(A.float(), W.float(), bias, (taps - 1) // 2, act, res, None, None, 0)
    if ln:
        want = torch.nn.functional.layer_norm(want, (N,), gamma.double(), beta.double(), 1e-5)
    if use_mask:
        want = want * (vpos < torch.minimum(torch.full_like(room, extra), room)).unsqueeze(1)

    d = lambda t: t.to(DEV) if t is not None else None
    dA, dW, db, dres, dg, dbe, dv, dr = map(d, (A, W, bias, res, gamma, beta, vpos, room))
    c32 = torch.full((rows, N), float("nan"), device=DEV) if out32 else None
    c16 = torch.full((rows, N), float("nan"), device=DEV, dtype=torch.bfloat16) if out16 else None
    code = lib().fs2_op_conv_gemm_bf16(stream(), ptr(dA), K, rows, ptr(dW), ptr(db), taps, (taps - 1) // 2, K, N, act,
                                       ptr(dres), N, ptr(dg), ptr(dbe), ptr(dv), ptr(dr), extra, ptr(c32), N, ptr(c16), N)
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    if out32:
        got = c32.cpu().double()
        assert torch.isfinite(got).all()
        err = (got - want).abs().max().item()
        assert err < 3e-4, f"fp32 output: max abs err {err}"
    if out16:
        got = c16.cpu().double()
        assert torch.isfinite(got).all()
        # one bf16 rounding of the exact value: relative 2^-9 (+ the fp32 accumulation noise)
        bound = want.abs() * 2.0 ** -8 + 4e-4
        assert ((got - want).abs() <= bound).all(), f"bf16 output: max abs err {(got - want).abs().max().item()}"
        if out32:   # the bf16 copy is the rounding of the fp32 result, bit for bit
            assert torch.equal(c16.cpu(), c32.cpu().bfloat16())


@pytest.mark.parametrize("name", ["c1_single", "pads", "controls", "longform"])
def test_bf16_forward_against_reference_fixture(name, sd32):
    model = model_for(sd32, math_mode="bf16")
    batch, kw, want, stride = load_golden(name)
    src_lens, mel_lens = batch["src_lens"].tolist(), want["mel_lens"].tolist()
    free = run(model, batch, **kw)
    for i, n in ((4, "log_d"), (2, "pitch")):
        mx, mean = err_stats(valid_rows(free[i].cpu().numpy(), src_lens), valid_rows(want[n], src_lens))
        log_diag(f"bf16 {name} {n}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= TOL_PRED, (name, n, mx)
    assert np.array_equal(free[6].cpu().numpy(), want["src_mask"])
    forced_p = run(model, batch, p_targets=torch.as_tensor(want["pitch"]).float(), **kw)
    mx, mean = err_stats(valid_rows(forced_p[3].cpu().numpy(), src_lens), valid_rows(want["energy"], src_lens))
    log_diag(f"bf16 {name} energy (pitch forced): max {mx:.3e} mean {mean:.3e}")
    assert mx <= TOL_PRED
    got = run(model, batch, **teacher_kwargs(want, src_lens))
    assert np.array_equal(got[9].cpu().numpy(), want["mel_lens"])      # integer stages: bit-exact
    assert np.array_equal(got[7].cpu().numpy(), want["mel_mask"])
    for i, n in ((0, "mel"), (1, "postnet")):
        gm = got[i].cpu().numpy()
        lens = mel_lens
        if stride > 1:
            gm = gm[:, ::stride]
            lens = [(int(l) + stride - 1) // stride for l in mel_lens]
        mx, mean = err_stats(valid_rows(gm, lens), valid_rows(want[n], lens))
        log_diag(f"bf16 {name} {n}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN, (name, n, mx, mean)
    mel = got[0].cpu().numpy()
    bias = sd32["mel_linear.bias"].numpy()
    for b, t in enumerate(mel_lens):
        if t < mel.shape[1]:
            assert np.array_equal(mel[b, t:], np.broadcast_to(bias, mel[b, t:].shape))


def test_bf16_close_to_tf32_at_full_size(sd32, syn):
    """Config 2 (batch 64): bf16 and TF32 modes agree within the bf16 budget when both are teacher-forced."""
    batch = syn.make_batch(syn.random_lengths(64, seed=2), seed=202)
    m32, m16 = model_for(sd32), model_for(sd32, math_mode="bf16")
    a = run(m32, batch)
    tk = dict(d_targets=a[5].cpu(), p_targets=a[2].cpu(), e_targets=a[3].cpu())
    a = run(m32, batch, **tk)
    b = run(m16, batch, **tk)
    assert torch.equal(a[9], b[9]) and torch.equal(a[7], b[7])
    lens = a[9].tolist()
    for i, n in ((0, "mel"), (1, "postnet")):
        mx, mean = err_stats(valid_rows(b[i].cpu().numpy(), lens), valid_rows(a[i].cpu().numpy(), lens))
        log_diag(f"bf16 vs tf32 config2 {n}: max {mx:.3e} mean {mean:.3e}")
        assert mx <= TOL_MEL_MAX and mean <= TOL_MEL_MEAN


@pytest.mark.parametrize("lens", [[1], [5, 64, 65, 33], [200, 7, 129, 128, 127], [700], [0, 3, 0, 300, 1], [37] * 70 + [513, 2, 1024]])
def test_attention_bf16(lens):
    """The bf16 attention kernel (128-key tiles, P as packed bf16 in tensor memory, V as the MN-major operand) against
    float64 softmax attention on the SAME bf16-rounded q / k / v: what remains is the bf16 rounding of the probabilities
    (2^-9 relative) and fp32 accumulation order."""
    g = torch.Generator().manual_seed(sum(lens) + 1)
    gap = 4
    starts, r = [], gap
    for n in lens:
        starts.append(r)
        r += n + gap
    rows = r
    qkv = torch.randn(rows, 768, generator=g).to(torch.bfloat16)
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
    code = lib().fs2_op_attention_bf16(stream(), ptr(dq), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
    assert code == 0, lib().fs2_last_error(None)
    torch.cuda.synchronize()
    got = out.cpu().double()
    live = torch.zeros(rows, dtype=torch.bool)
    for s, n in zip(starts, lens):
        live[s: s + n] = True
    assert (got[~live] == 0).all(), "rows outside the utterances must not be written"
    err = (got[live] - want[live]).abs().max().item()
    assert torch.isfinite(got).all()
    print(f"bf16 attention lens={lens[:4]}.. max abs err {err:.3e}")
    assert err < 1.5e-2, f"max abs err {err}"   # unit-variance data: |p v| sums of up to 1024 terms with 2^-9 relative weights
