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
