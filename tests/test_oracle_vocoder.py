"""The HiFi-GAN oracle (oracle/hifigan_oracle.py) against the output of the unmodified reference generator
(tests/golden/vocoder.npz, made by tests/golden/make_golden_vocoder.py) -- the pinning step for SURVEY.md §8(f) rank 2."""
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as H
from helpers import GOLDEN_DIR


@pytest.fixture(scope="module")
def vsd(syn):
    return syn.synthetic_vocoder_state_dict(seed=0)


def test_schema_has_the_generators_156_tensors(syn, vsd):
    z = np.load(os.path.join(GOLDEN_DIR, "vocoder.npz"))
    assert len(vsd) == int(z["n_keys"]) == 156
    assert [k for k, _ in syn.vocoder_schema()] == list(vsd)


def test_oracle_fp64_equals_reference_generator(vsd):
    z = np.load(os.path.join(GOLDEN_DIR, "vocoder.npz"))
    sd64 = {k: v.double() for k, v in vsd.items()}
    got = H.generator(sd64, torch.from_numpy(z["mel"])).numpy()
    assert got.shape == z["wav"].shape == (2, 1, 21 * 256)
    assert np.max(np.abs(got - z["wav"])) <= 1e-12
    assert 0.1 < float(z["rms"]) < 0.7, "fixture must not sit in tanh saturation"


def test_weight_norm_checkpoints_fold_to_plain_weights(vsd):
    """A checkpoint in weight_norm form (weight_g / weight_v) gives the same generator (models.py:169-174)."""
    wn = {}
    for k, v in vsd.items():
        if k.endswith(".weight") and not k.startswith("conv_post"):
            base = k[: -len(".weight")]
            g = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            wn[base + ".weight_g"], wn[base + ".weight_v"] = g, v * 3.0      # any positive rescaling of v is absorbed
        else:
            wn[k] = v
    folded = H.fold_weight_norm(wn)
    assert set(folded) == set(vsd)
    for k in vsd:
        assert torch.allclose(folded[k], vsd[k], rtol=1e-5, atol=1e-7), k


@pytest.mark.skipif(not os.path.isdir("/root/reference/hifigan"), reason="reference tree only exists in the authoring container")
def test_oracle_against_live_reference_generator(vsd):
    import sys
    sys.path.insert(0, GOLDEN_DIR)
    import make_golden_vocoder
    gen = make_golden_vocoder.build_reference_generator(vsd, torch.float32)
    mel = torch.randn(1, 80, 9, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = gen(mel)
    got = H.generator(vsd, mel)
    assert float((got - want).abs().max()) <= 1e-5


def test_vocoder_infer_int16_and_trim(vsd):
    mel = torch.randn(2, 80, 5, generator=torch.Generator().manual_seed(2))
    wavs = H.vocoder_infer(vsd, mel, lengths=[5 * 256, 3 * 256])
    assert wavs[0].dtype == np.int16 and wavs[0].shape == (1280,) and wavs[1].shape == (768,)


def test_facade_state_dict_is_the_generators_and_folds_weight_norm(syn, vsd):
    """HiFiGANGeneratorB200 holds the reference generator's 156 tensors (remove_weight_norm() form), loads a weight_norm
    checkpoint (utils/model.py:60-66) and has no CPU path."""
    import fs2_b200
    voc = fs2_b200.HiFiGANGeneratorB200()
    assert [(k, tuple(v.shape)) for k, v in voc.state_dict().items()] == [(k, tuple(s)) for k, s in syn.vocoder_schema()]
    wn = {}
    for k, v in vsd.items():
        if k.endswith(".weight") and not k.startswith("conv_post"):
            base = k[: -len(".weight")]
            wn[base + ".weight_g"] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            wn[base + ".weight_v"] = v * 0.5
        else:
            wn[k] = v
    res = voc.load_state_dict(wn, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in vsd.items():
        assert torch.allclose(voc.state_dict()[k], v, rtol=1e-5, atol=1e-7), k
    with pytest.raises(RuntimeError, match="CUDA"):
        voc(torch.zeros(1, 80, 4))
    with pytest.raises(RuntimeError, match="inference"):
        voc.train()
    with pytest.raises(ValueError):
        fs2_b200.HiFiGANGeneratorB200({"upsample_rates": [8, 8, 4]})
