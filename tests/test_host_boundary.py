"""Host-side boundary checks that need no GPU: the facade's parameter tree is the reference's
state dict, the library loads and exports every symbol of include/fs2_b200.h, and the product
refuses to run without a CUDA device (no CPU path)."""
import ctypes
import os
import re
import tempfile

import pytest
import torch

import fs2_b200
from fs2_b200 import _lib, build
from helpers import GOLDEN_DIR

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def facade(syn):
    d = syn.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    return fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(d), fs2_b200.config.default_model_config())


@pytest.fixture(scope="module")
def library():
    build.build_library()
    return _lib.load_library()


def test_facade_state_dict_is_the_references(facade):
    want = {}
    for line in open(os.path.join(GOLDEN_DIR, "state_dict_keys.txt")):
        key, rest = line.split(" ", 1)
        shape, dtype = rest.rsplit(" ", 1)
        want[key] = (eval(shape), dtype.strip())
    got = facade.state_dict()
    assert set(got) == set(want)
    assert len(got) == 240
    for k, (shape, dtype) in want.items():
        assert tuple(got[k].shape) == shape, k
        assert str(got[k].dtype).replace("torch.", "") == dtype, k


def test_reference_checkpoint_loads_strict(facade, sd32):
    res = facade.load_state_dict(sd32, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(facade.state_dict()["postnet.convolutions.3.1.running_var"], sd32["postnet.convolutions.3.1.running_var"])


def test_no_cpu_path_and_no_training(facade, syn):
    b = syn.config1_batch()
    with pytest.raises(RuntimeError, match="CUDA"):
        facade(b["speakers"], b["emotions"], b["arousals"], b["valences"], b["texts"], b["src_lens"], b["max_src_len"])
    with pytest.raises(RuntimeError, match="inference"):
        facade.train()


def test_unsupported_hyperparameters_are_rejected(syn):
    d = syn.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
    cfg = fs2_b200.config.default_model_config()
    cfg["transformer"]["encoder_hidden"] = 384
    with pytest.raises(ValueError):
        fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(d), cfg)


def test_library_exports_every_declared_symbol(library):
    header = open(os.path.join(REPO, "include", "fs2_b200.h")).read()
    declared = set(re.findall(r"\b(fs2_[a-z0-9_]+)\s*\(", header))
    declared -= {"fs2_ctx"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(library, name) is not None
    assert library.fs2_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_create_fails_loudly_without_a_gpu(library):
    cfg = _lib.Config(n_src_vocab=139, n_speaker=10, n_emotion=5, n_arousal=4, n_valence=5, max_seq_len=2000,
                      math_mode=0)
    ctx = ctypes.c_void_p()
    code = library.fs2_create(ctypes.byref(cfg), 0, ctypes.byref(ctx))
    assert code != 0
    assert library.fs2_last_error(None)
