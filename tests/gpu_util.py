"""Helpers for the -m gpu tests: everything calls the product through the C ABI
(fs2_b200._lib ctypes binding or the FastSpeech2B200 facade)."""
import ctypes as C
import os
import tempfile

import numpy as np
import torch

import fs2_b200
from fs2_b200 import _lib, build

DEV = "cuda:0"
_MODELS = {}


def lib():
    build.build_library()
    return _lib.load_library()


def model_for(sd, math_mode="tf32", key="seed0", pitch_level="phoneme_level", energy_level="phoneme_level"):
    k = (key, math_mode, pitch_level, energy_level)
    if k not in _MODELS:
        d = fs2_b200.synthetic.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))
        pre = fs2_b200.config.default_preprocess_config(d)
        pre["preprocessing"]["pitch"]["feature"] = pitch_level
        pre["preprocessing"]["energy"]["feature"] = energy_level
        m = fs2_b200.FastSpeech2B200(pre,
                                     fs2_b200.config.default_model_config(), math_mode=math_mode)
        m.load_state_dict(sd)
        _MODELS[k] = m.to(DEV)
    return _MODELS[k]


def to_dev(batch):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}


def run(model, batch, **kw):
    b = to_dev(batch)
    kw = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in kw.items()}
    out = model(b["speakers"], b["emotions"], b["arousals"], b["valences"], b["texts"], b["src_lens"],
                b["max_src_len"], **kw)
    torch.cuda.synchronize()
    return out


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def round_tf32(x):
    """cvt.rna.tf32.f32 on the host: round to 10 explicit mantissa bits, ties away from zero."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def packed_to_padded(tap, starts, lens, max_len):
    """[rows, C] packed rows -> [B, max_len, C] with zeros on padding."""
    B = len(lens)
    out = np.zeros((B, max_len, tap.shape[1]), dtype=tap.dtype)
    for b in range(B):
        out[b, : lens[b]] = tap[starts[b]: starts[b] + lens[b]]
    return out


def err_stats(got, want):
    d = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64))
    return float(d.max(initial=0.0)), float(d.mean()) if d.size else 0.0
