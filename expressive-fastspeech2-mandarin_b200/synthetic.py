"""Deterministic synthetic weights, JSON fixtures and input batches.

There are no trained checkpoints (Git-LFS stubs) and no preprocessed_data/ in the
reference tree, so every parity test, smoke() and bench.py runs on weights made
here.  Each tensor is drawn from its own CPU generator seeded by crc32(key)^seed,
so the same 240-key state dict (schema: SURVEY.md Appendix A.1) can be rebuilt on
any box and loaded into BOTH the reference module and `FastSpeech2B200`.
"""
import json
import math
import os
import zlib

import numpy as np
import torch

from .config import N_SRC_VOCAB, POSTNET_DIM, POSTNET_KERNEL, POSTNET_LAYERS

SPEAKERS = {f"{i:04d}": i - 1 for i in range(1, 11)}
EMOTIONS = {
    "emotion_dict": {"Angry": 0, "Happy": 1, "Neutral": 2, "Sad": 3, "Surprise": 4},
    "arousal_dict": {"0.9": 0, "0.8": 1, "0.5": 2, "0.3": 3},
    "valence_dict": {"0.1": 0, "0.8": 1, "0.5": 2, "0.2": 3, "0.6": 4},
}
STATS = {"pitch": [-2.5, 9.0, 0.0, 1.0], "energy": [-1.5, 8.0, 0.0, 1.0]}

# "j i n t ia n t ia n q i zh e n h ao" under text/symbols_pinyin.py (last duplicate key wins).
C1_IDS = [82, 77, 86, 96, 78, 86, 96, 78, 86, 91, 77, 107, 71, 86, 76, 66]


def write_fixture_jsons(directory):
    """speakers.json / emotions.json / stats.json in the formats written by
    preprocessor/preprocessor.py:183-205 and read by model/fastspeech2.py:31-54,
    model/modules.py:41-46."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "speakers.json"), "w") as f:
        json.dump(SPEAKERS, f)
    with open(os.path.join(directory, "emotions.json"), "w") as f:
        json.dump(EMOTIONS, f)
    with open(os.path.join(directory, "stats.json"), "w") as f:
        json.dump(STATS, f)
    return directory


def sinusoid_table(n_position, d_hid=256):
    """Vectorised float64 evaluation of transformer/Models.py:10-30 (bitwise equal to
    the reference's Python double loop), cast to fp32 like torch.FloatTensor does."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    denom = np.power(10000.0, 2.0 * (j // 2) / d_hid)
    table = pos / denom[None, :]
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table.astype(np.float32))


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(key, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32) * (hi - lo) + lo


def _normal(key, seed, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(key, seed), dtype=torch.float32) * std + mean


def synthetic_state_dict(seed=0, duration_bias=1.9, model_config=None):
    """All 240 tensors of FastSpeech2.state_dict().

    Dense layers use U(-1/sqrt(fan_in), +1/sqrt(fan_in)) like torch's defaults;
    LayerNorm affine and BatchNorm statistics are randomised so that LN affine and
    BN folding are actually exercised; the duration head bias is raised so that
    durations are realistic (about 4-6 frames per phoneme instead of 0).
    """
    sd = {}

    def dense(prefix, out_c, in_c, k=None):
        fan_in = in_c * (k or 1)
        b = 1.0 / math.sqrt(fan_in)
        shape = (out_c, in_c) if k is None else (out_c, in_c, k)
        sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, shape, -b, b)
        sd[prefix + ".bias"] = _uniform(prefix + ".bias", seed, (out_c,), -b, b)

    def layer_norm(prefix, n):
        sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, (n,), 0.8, 1.2)
        sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (n,), 0.05)

    d, dk_total, d_inner = 256, 256, 1024
    pe = sinusoid_table(2001, d).unsqueeze(0)
    sd["encoder.src_word_emb.weight"] = _normal("encoder.src_word_emb.weight", seed, (N_SRC_VOCAB, d), 1.0)
    sd["encoder.position_enc"] = pe.clone()
    for stack, n_layers in (("encoder", 4), ("decoder", 6)):
        if stack == "decoder":
            sd["decoder.position_enc"] = pe.clone()
        for i in range(n_layers):
            p = f"{stack}.layer_stack.{i}"
            for w in ("w_qs", "w_ks", "w_vs"):
                dense(f"{p}.slf_attn.{w}", dk_total, d)
            layer_norm(f"{p}.slf_attn.layer_norm", d)
            dense(f"{p}.slf_attn.fc", d, dk_total)
            dense(f"{p}.pos_ffn.w_1", d_inner, d, 9)
            dense(f"{p}.pos_ffn.w_2", d, d_inner, 1)
            layer_norm(f"{p}.pos_ffn.layer_norm", d)

    va = "variance_adaptor"
    for name in ("duration", "pitch", "energy"):
        p = f"{va}.{name}_predictor"
        dense(f"{p}.conv_layer.conv1d_1.conv", 256, 256, 3)
        layer_norm(f"{p}.conv_layer.layer_norm_1", 256)
        dense(f"{p}.conv_layer.conv1d_2.conv", 256, 256, 3)
        layer_norm(f"{p}.conv_layer.layer_norm_2", 256)
        dense(f"{p}.linear_layer", 1, 256)
    sd[f"{va}.duration_predictor.linear_layer.bias"] = torch.tensor([duration_bias], dtype=torch.float32)
    sd[f"{va}.pitch_bins"] = torch.linspace(STATS["pitch"][0], STATS["pitch"][1], 255)
    sd[f"{va}.energy_bins"] = torch.linspace(STATS["energy"][0], STATS["energy"][1], 255)
    sd[f"{va}.pitch_embedding.weight"] = _normal(f"{va}.pitch_embedding.weight", seed, (256, 256), 1.0)
    sd[f"{va}.energy_embedding.weight"] = _normal(f"{va}.energy_embedding.weight", seed, (256, 256), 1.0)

    dense("mel_linear", 80, 256)
    chans = [80] + [POSTNET_DIM] * (POSTNET_LAYERS - 1) + [80]
    for j in range(POSTNET_LAYERS):
        dense(f"postnet.convolutions.{j}.0.conv", chans[j + 1], chans[j], POSTNET_KERNEL)
        bn = f"postnet.convolutions.{j}.1"
        sd[bn + ".weight"] = _uniform(bn + ".weight", seed, (chans[j + 1],), 0.5, 1.5)
        sd[bn + ".bias"] = _normal(bn + ".bias", seed, (chans[j + 1],), 0.1)
        sd[bn + ".running_mean"] = _normal(bn + ".running_mean", seed, (chans[j + 1],), 0.1)
        sd[bn + ".running_var"] = _uniform(bn + ".running_var", seed, (chans[j + 1],), 0.5, 1.5)
        sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    sd["speaker_emb.weight"] = _normal("speaker_emb.weight", seed, (len(SPEAKERS), 256), 1.0)
    sd["emotion_emb.weight"] = _normal("emotion_emb.weight", seed, (len(EMOTIONS["emotion_dict"]), 128), 1.0)
    sd["arousal_emb.weight"] = _normal("arousal_emb.weight", seed, (len(EMOTIONS["arousal_dict"]), 64), 1.0)
    sd["valence_emb.weight"] = _normal("valence_emb.weight", seed, (len(EMOTIONS["valence_dict"]), 64), 1.0)
    dense("emotion_linear.0", 256, 256)
    return sd


def state_dict_checksum(sd):
    """A float64 fingerprint used by the golden fixtures to detect RNG drift."""
    acc = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        acc += float(t.sum()) + 0.5 * float((t * t).sum())
    return acc


# ----------------------------------------------------------------------------------------
# Input batches.  Every batch is the dict of arguments of
# FastSpeech2.forward (model/fastspeech2.py:73-91), as CPU int64 tensors + a host int.
# ----------------------------------------------------------------------------------------

def make_batch(lengths, seed=0, ids=None):
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    lengths = [int(x) for x in lengths]
    B, L = len(lengths), max(lengths)
    texts = torch.zeros(B, L, dtype=torch.int64)
    for b, n in enumerate(lengths):
        if ids is not None:
            texts[b, :n] = torch.tensor(ids[b], dtype=torch.int64)
        else:
            texts[b, :n] = torch.randint(1, 108, (n,), generator=g)
    return {
        "speakers": torch.randint(0, len(SPEAKERS), (B,), generator=g),
        "emotions": torch.randint(0, 5, (B,), generator=g),
        "arousals": torch.randint(0, 4, (B,), generator=g),
        "valences": torch.randint(0, 5, (B,), generator=g),
        "texts": texts,
        "src_lens": torch.tensor(lengths, dtype=torch.int64),
        "max_src_len": L,
    }


def config1_batch():
    """BASELINE config 1: '今天天气真好', speaker 0001, Happy (arousal 0.8, valence 0.8)."""
    b = make_batch([len(C1_IDS)], ids=[C1_IDS])
    b["speakers"] = torch.tensor([0])
    b["emotions"] = torch.tensor([1])
    b["arousals"] = torch.tensor([1])
    b["valences"] = torch.tensor([1])
    return b


def random_lengths(batch, lo=20, hi=120, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    return torch.randint(lo, hi + 1, (batch,), generator=g).tolist()


def config2_batch(seed=0, batch=64):
    """BASELINE config 2: batch 64, 20-120 phonemes, mixed speakers/emotions."""
    return make_batch(random_lengths(batch, seed=seed), seed=seed)


def config3_batch(seed=0, batch=512):
    return make_batch(random_lengths(batch, seed=seed + 100), seed=seed + 100)


def config4_batch(seed=0, length=400):
    return make_batch([length], seed=seed + 200)


def config5_batch(seed=0, batch=32):
    return make_batch(random_lengths(batch, seed=seed + 300), seed=seed + 300)


def algorithmic_flops(src_lens, mel_lens):
    """Valid-row FLOPs of one forward (MAC = 2), SURVEY.md Appendix C / BASELINE.md §4."""
    total = 0
    for L, T in zip(src_lens, mel_lens):
        L, T = int(L), int(T)
        total += L * 25_429_504 + 4096 * L * L + 131_072 + T * 43_327_488 + 6144 * T * T
    return total


# ---------------------------------------------------------------------------------------------------
# HiFi-GAN generator (hifigan/models.py:112-146, hifigan/config.json): the reference ships no vocoder
# weights either (.MISSING_LARGE_BLOBS), so parity runs on seeded weights of the V1 architecture.
VOC_UP_RATES, VOC_UP_KERNELS, VOC_UP_INITIAL = (8, 8, 2, 2), (16, 16, 4, 4), 512
VOC_RES_KERNELS, VOC_RES_DILATIONS = (3, 7, 11), (1, 3, 5)


def vocoder_schema():
    """(key, shape) of Generator.state_dict() after remove_weight_norm(), in its order."""
    out = [("conv_pre.weight", (VOC_UP_INITIAL, 80, 7)), ("conv_pre.bias", (VOC_UP_INITIAL,))]
    ch = VOC_UP_INITIAL
    for i, k in enumerate(VOC_UP_KERNELS):
        out += [(f"ups.{i}.weight", (ch, ch // 2, k)), (f"ups.{i}.bias", (ch // 2,))]   # ConvTranspose1d: [in, out, k]
        ch //= 2
    ch = VOC_UP_INITIAL
    for i in range(len(VOC_UP_RATES)):
        ch //= 2
        for j, k in enumerate(VOC_RES_KERNELS):
            n = i * len(VOC_RES_KERNELS) + j
            for grp in ("convs1", "convs2"):
                for m in range(3):
                    out += [(f"resblocks.{n}.{grp}.{m}.weight", (ch, ch, k)), (f"resblocks.{n}.{grp}.{m}.bias", (ch,))]
    out += [("conv_post.weight", (1, ch, 7)), ("conv_post.bias", (1,))]
    return out


def synthetic_vocoder_state_dict(seed=0):
    """Variance-preserving seeded weights (the reference's N(0, 0.01) init, models.py:10-13, would make every
    activation ~1e-6 and the parity test vacuous): std = gain / sqrt(fan_in taps that actually contribute)."""
    sd = {}
    for key, shape in vocoder_schema():
        g = _gen("voc." + key, seed)
        if key.endswith("bias"):
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        elif key.startswith("ups."):
            cin, _, k = shape
            stride = VOC_UP_RATES[int(key.split(".")[1])]
            sd[key] = torch.randn(shape, generator=g) * (1.3 / math.sqrt(cin * k / stride))
        else:
            _, cin, k = shape
            gain = 0.6 if ".convs2." in key else 1.3      # the residual branch adds to x: keep the sum bounded
            if key.startswith("conv_pre") or key.startswith("conv_post"):
                gain = 0.45                               # activations ~1, pre-tanh output ~0.4 (tanh not saturated)
            sd[key] = torch.randn(shape, generator=g) * (gain / math.sqrt(cin * k))
    return sd
