"""HiFi-GAN generator on the B200 engine through the C ABI (fs2_voc_*), against
(1) the committed output of the unmodified reference generator (tests/golden/vocoder.npz) and (2) the CPU oracle.

Stated tolerance (TF32 operands, fp32 accumulation, 50 convolutions deep): waveform max-abs <= 4e-3, mean-abs <= 5e-4
on samples in [-1, 1] (measured on B200: 1.3e-3 / 1.9e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as H
from gpu_util import DEV
from helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL_MAX, TOL_MEAN = 4e-3, 5e-4


@pytest.fixture(scope="module")
def vsd(syn):
    return syn.synthetic_vocoder_state_dict(seed=0)


@pytest.fixture(scope="module")
def voc(vsd):
    import fs2_b200
    m = fs2_b200.HiFiGANGeneratorB200()
    m.load_state_dict(vsd)
    return m.to(DEV)


def check(got, want, tag):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    print(f"vocoder {tag}: max {d.max():.3e} mean {d.mean():.3e} (rms of the signal {np.sqrt((want ** 2).mean()):.3f})")
    assert np.isfinite(got).all()
    assert d.max() <= TOL_MAX and d.mean() <= TOL_MEAN, (tag, d.max(), d.mean())


def test_reference_fixture(voc):
    z = np.load(os.path.join(GOLDEN_DIR, "vocoder.npz"))
    wav = voc(torch.from_numpy(z["mel"]).float().to(DEV))
    torch.cuda.synchronize()
    assert tuple(wav.shape) == z["wav"].shape
    check(wav.cpu().numpy(), z["wav"], "reference fixture")


@pytest.mark.parametrize("shape", [(1, 1), (1, 37), (3, 50), (2, 130)])
def test_against_oracle(voc, vsd, shape):
    B, T = shape
    mel = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(B * 100 + T)) * 1.5 - 2.0
    want = H.generator({k: v.double() for k, v in vsd.items()}, mel.double()).numpy()
    got = voc(mel.to(DEV))
    torch.cuda.synchronize()
    check(got.cpu().numpy(), want, f"B{B} T{T}")


def test_transposed_input_and_lengths(voc, vsd):
    """[B, T, 80].transpose(1, 2) goes in without a copy; with mel_lens every utterance equals its solo synthesis."""
    lens = [40, 17, 33]
    mel_bt = torch.randn(3, 40, 80, generator=torch.Generator().manual_seed(9)) * 1.5 - 2.0
    got = voc(mel_bt.to(DEV).transpose(1, 2), mel_lens=torch.tensor(lens))
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    sd64 = {k: v.double() for k, v in vsd.items()}
    for b, n in enumerate(lens):
        want = H.generator(sd64, mel_bt[b: b + 1, :n].transpose(1, 2).double()).numpy()[0, 0]
        check(got[b, 0, : n * 256], want, f"solo semantics utt {b}")
        assert (got[b, 0, n * 256:] == 0).all()


def test_vocoder_infer_matches_reference_api(voc, vsd):
    import fs2_b200
    mel = torch.randn(2, 80, 12, generator=torch.Generator().manual_seed(4)) * 1.5 - 2.0
    pre = fs2_b200.config.default_preprocess_config()
    wavs = fs2_b200.vocoder_infer(mel.to(DEV), voc, fs2_b200.config.default_model_config(), pre, lengths=[12 * 256, 9 * 256])
    want = H.vocoder_infer(vsd, mel[:, :, :12], lengths=[12 * 256, 9 * 256])
    assert wavs[0].dtype == np.int16 and wavs[0].shape == (12 * 256,) and wavs[1].shape == (9 * 256,)
    assert np.abs(wavs[0].astype(np.int32) - want[0].astype(np.int32)).max() <= 132     # 4e-3 of full scale


def test_cpu_tensor_is_rejected(voc):
    with pytest.raises(RuntimeError):
        voc(torch.zeros(1, 80, 4))


# ------------------------------------------------------------------ the vocoder's GEMM form, operator level
def _conv_ex_ref(A, W, bias, dil, act, slope, res, inv, act2, live):
    rows, K = A.shape
    taps, N, _ = W.shape
    pad = dil * (taps - 1) // 2
    A64 = torch.zeros(rows + 2 * pad + 1, K, dtype=torch.float64)
    A64[pad: pad + rows] = A.double()
    out = bias.double().unsqueeze(0).repeat(rows, 1)
    for t in range(taps):
        out += A64[t * dil: t * dil + rows] @ W[t].double().T
    lrelu = lambda x: torch.where(x >= 0, x, x * slope)
    if act == 3:
        out = lrelu(out)
    if res is not None:
        r = res.double()
        out = out + (torch.where(r >= 0, r, r / slope) if inv else r)
    if act2 == 3:
        out = lrelu(out)
    return out * live.unsqueeze(1)


EX_CASES = [
    # rows, K, N, taps, dil, residual, mask_shift, name             (K <= 64 with several taps -> A-resident variant)
    (1000, 32, 32, 3, 1, False, 0, "c32_k3"),
    (5000, 32, 32, 11, 5, True, 3, "c32_k11_d5_res"),
    (3001, 64, 64, 7, 3, True, 1, "c64_k7_d3_res"),
    (2500, 64, 64, 11, 5, False, 0, "c64_k11_d5"),
    (777, 64, 64, 3, 1, True, 0, "up_64_to_2x32"),
    (900, 128, 128, 11, 5, True, 2, "c128_k11_d5_streaming"),
    (300, 256, 256, 7, 3, True, 0, "c256_k7_d3_streaming"),
    (40000, 32, 32, 7, 5, True, 8, "c32_many_tiles"),
]


@pytest.mark.parametrize("resident", [1, 0])
@pytest.mark.parametrize("case", EX_CASES, ids=[c[-1] for c in EX_CASES])
def test_conv_gemm_ex(case, resident):
    from gpu_util import lib, ptr, round_tf32, stream
    rows, K, N, taps, dil, use_res, shift, _ = case
    g = torch.Generator().manual_seed(rows + K + taps * 13 + dil)
    A = round_tf32(torch.randn(rows, K, generator=g))
    W = round_tf32(torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps))
    bias = torch.randn(N, generator=g) * 0.2
    res = torch.randn(rows, N, generator=g) if use_res else None
    n_mask = (rows >> shift) + 1
    vpos = torch.randint(-6, 2, (n_mask,), generator=g, dtype=torch.int32)
    room = torch.zeros(n_mask, dtype=torch.int32)
    live = (vpos < 0)[torch.arange(rows) >> shift]
    slope = 0.1
    want = _conv_ex_ref(A, W, bias, dil, 3 if not use_res else 0, slope, res, True, 3 if use_res else 0, live)
    d = lambda t: t.to(DEV) if t is not None else None
    dA, dW, db, dres, dv, dr = map(d, (A, W, bias, res, vpos, room))
    out = torch.full((rows, N), float("nan"), device=DEV)
    L = lib()
    L.fs2_debug_set_flag(3, resident)
    try:
        code = L.fs2_op_conv_gemm_ex(stream(), ptr(dA), K, rows, ptr(dW), ptr(db), taps, dil, K, N, 0 if use_res else 3, slope,
                                     ptr(dres), N, 1 if use_res else 0, 3 if use_res else 0, ptr(dv), ptr(dr), 0, shift,
                                     ptr(out), N)
        assert code == 0, L.fs2_last_error(None)
        torch.cuda.synchronize()
    finally:
        L.fs2_debug_set_flag(3, 1)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err < 3e-4, f"max abs err {err}"


# ------------------------------------------------------------------ bf16 mode of the vocoder
def test_bf16_vocoder_against_reference_fixture_and_oracle(vsd):
    """FS2_MATH_BF16 vocoder: bf16 weights and activations (fp32 accumulation, biases, output).  Stated tolerance on the
    waveform in [-1, 1], ~50 convolutions deep with the ResBlock residual chains held in bf16: max-abs <= 3e-2,
    mean-abs <= 4e-3 (measured on B200: 1.1e-2 / 1.9e-3)."""
    import fs2_b200
    voc16 = fs2_b200.HiFiGANGeneratorB200(math_mode="bf16")
    voc16.load_state_dict(vsd)
    voc16 = voc16.to(DEV)
    z = np.load(os.path.join(GOLDEN_DIR, "vocoder.npz"))
    got = voc16(torch.from_numpy(z["mel"]).float().to(DEV)).cpu().numpy()
    d = np.abs(got.astype(np.float64) - z["wav"])
    print(f"bf16 vocoder reference fixture: max {d.max():.3e} mean {d.mean():.3e}")
    assert np.isfinite(got).all() and d.max() <= 3e-2 and d.mean() <= 4e-3
    lens = [40, 17, 33]
    mel_bt = torch.randn(3, 40, 80, generator=torch.Generator().manual_seed(9)) * 1.5 - 2.0
    got = voc16(mel_bt.to(DEV).transpose(1, 2), mel_lens=torch.tensor(lens)).cpu().numpy()
    sd64 = {k: v.double() for k, v in vsd.items()}
    for b, n in enumerate(lens):
        want = H.generator(sd64, mel_bt[b: b + 1, :n].transpose(1, 2).double()).numpy()[0, 0]
        d = np.abs(got[b, 0, : n * 256] - want)
        print(f"bf16 vocoder utt {b}: max {d.max():.3e} mean {d.mean():.3e}")
        assert d.max() <= 3e-2 and d.mean() <= 4e-3
        assert (got[b, 0, n * 256:] == 0).all()
