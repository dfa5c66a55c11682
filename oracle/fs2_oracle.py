"""CPU ORACLE for the FastSpeech2 inference forward pass  --  TEST INFRASTRUCTURE ONLY.

This file restates, as flat functions over a state dict, the algorithm of the
reference's `FastSpeech2.forward` (all citations are /root/reference paths).  It is
imported ONLY by tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` leg, as the checker or the timed CPU baseline -- never by the
product package (`expressive-fastspeech2-mandarin_b200/`), which has no CPU path.

Pinning: the reference ships no golden vectors or tests for this path,
and its arithmetic lives in PyTorch (requirements.txt pins no version; this image has
torch 2.11.0).  The oracle is therefore pinned against outputs of the reference itself,
generated in the authoring container by tests/golden/make_golden.py (which imports
/root/reference, loads the same synthetic state dict with strict=True and runs the
unmodified module in float64) and committed under tests/golden/*.npz;
tests/test_oracle_golden.py checks this file against them (fp64: <= 1e-12).

The arithmetic uses the same torch CPU operators the reference calls (F.conv1d,
F.linear, torch.bmm, F.layer_norm, F.batch_norm, torch.bucketize), in the same order,
so fp32 results agree with the reference to rounding and the CPU timing of this port is
representative of the reference's own CPU forward.
"""
import numpy as np
import torch
import torch.nn.functional as F

N_ENC_LAYERS = 4      # config/ESD-Chinese-Singing-MFA/model.yaml:2
N_DEC_LAYERS = 6      # model.yaml:5
N_HEAD = 2            # model.yaml:3,6
MAX_SEQ_LEN = 2000    # model.yaml:27
N_POSTNET = 5         # transformer/Layers.py:77


# ------------------------------------------------------------------ helpers

def pad_mask(lengths, max_len=None):
    """utils/tools.py:152-160 -- True marks padding."""
    if max_len is None:
        max_len = int(lengths.max().item())
    pos = torch.arange(0, int(max_len)).unsqueeze(0)
    return pos >= lengths.unsqueeze(1)


def sinusoid_rows(n_position, d_hid):
    """transformer/Models.py:10-30, evaluated vectorised in float64 (bit-identical to the
    reference's Python loops) and cast to fp32 as torch.FloatTensor does."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000.0, 2.0 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang.astype(np.float32))


def _position_rows(sd, key, n_rows, dtype):
    """Encoder: Models.py:82-91; decoder: Models.py:145-162 (eval mode never truncates)."""
    if n_rows > MAX_SEQ_LEN:
        return sinusoid_rows(n_rows, sd[key].shape[-1]).to(dtype)
    return sd[key][0, :n_rows, :]


# ------------------------------------------------------------------ FFT block

def attention(sd, p, x, key_pad):
    """transformer/SubLayers.py:29-57 + transformer/Modules.py:14-25."""
    B, S, D = x.shape
    dk = D // N_HEAD

    def heads(name):
        y = F.linear(x, sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"])
        return y.view(B, S, N_HEAD, dk).permute(2, 0, 1, 3).contiguous().view(N_HEAD * B, S, dk)

    q, k, v = heads("w_qs"), heads("w_ks"), heads("w_vs")
    scores = torch.bmm(q, k.transpose(1, 2)) / float(np.power(dk, 0.5))
    kmask = key_pad.unsqueeze(1).expand(-1, S, -1).repeat(N_HEAD, 1, 1)
    scores = scores.masked_fill(kmask, -np.inf)
    out = torch.bmm(torch.softmax(scores, dim=2), v)
    out = out.view(N_HEAD, B, S, dk).permute(1, 2, 0, 3).contiguous().view(B, S, D)
    out = F.linear(out, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"])
    return F.layer_norm(out + x, (D,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


def conv_ffn(sd, p, x):
    """transformer/SubLayers.py:85-93 -- Conv1d(k=9,pad=4) -> ReLU -> Conv1d(k=1), post-LN."""
    D = x.shape[-1]
    w1 = sd[f"{p}.w_1.weight"]
    h = F.conv1d(x.transpose(1, 2), w1, sd[f"{p}.w_1.bias"], padding=(w1.shape[-1] - 1) // 2)
    w2 = sd[f"{p}.w_2.weight"]
    h = F.conv1d(F.relu(h), w2, sd[f"{p}.w_2.bias"], padding=(w2.shape[-1] - 1) // 2)
    return F.layer_norm(h.transpose(1, 2) + x, (D,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


def fft_block(sd, p, x, pad):
    """transformer/Layers.py:21-30."""
    x = attention(sd, p + ".slf_attn", x, pad).masked_fill(pad.unsqueeze(-1), 0)
    return conv_ffn(sd, p + ".pos_ffn", x).masked_fill(pad.unsqueeze(-1), 0)


def encoder(sd, texts, pad, taps=None):
    """transformer/Models.py:73-100."""
    emb = sd["encoder.src_word_emb.weight"]
    L = texts.shape[1]
    x = F.embedding(texts, emb) + _position_rows(sd, "encoder.position_enc", L, emb.dtype).unsqueeze(0)
    if taps is not None:
        taps["enc_in"] = x
    for i in range(N_ENC_LAYERS):
        x = fft_block(sd, f"encoder.layer_stack.{i}", x, pad)
        if taps is not None:
            taps[f"enc_{i}"] = x
    return x


def decoder(sd, x, pad, taps=None):
    """transformer/Models.py:139-171 (eval)."""
    T = x.shape[1]
    x = x + _position_rows(sd, "decoder.position_enc", T, x.dtype).unsqueeze(0)
    if taps is not None:
        taps["dec_in"] = x
    for i in range(N_DEC_LAYERS):
        x = fft_block(sd, f"decoder.layer_stack.{i}", x, pad)
        if taps is not None:
            taps[f"dec_{i}"] = x
    return x


# ------------------------------------------------------------------ variance adaptor

def conditioning(sd, speakers, emotions, arousals, valences):
    """model/fastspeech2.py:101-110 -- the two per-utterance vectors added to every row."""
    spk = F.embedding(speakers, sd["speaker_emb.weight"])
    emo = torch.cat((F.embedding(emotions, sd["emotion_emb.weight"]),
                     F.embedding(arousals, sd["arousal_emb.weight"]),
                     F.embedding(valences, sd["valence_emb.weight"])), dim=-1)
    emo = F.relu(F.linear(emo, sd["emotion_linear.0.weight"], sd["emotion_linear.0.bias"]))
    return spk, emo


def variance_predictor(sd, p, x, pad):
    """model/modules.py:242-250 (layers :209-240; Conv wrapper :291-296).  The input is NOT masked."""
    h = x
    for n in (1, 2):
        w = sd[f"{p}.conv_layer.conv1d_{n}.conv.weight"]
        padding = (w.shape[-1] - 1) // 2 if n == 1 else 1          # modules.py:221 vs :230
        h = F.conv1d(h.transpose(1, 2), w, sd[f"{p}.conv_layer.conv1d_{n}.conv.bias"], padding=padding).transpose(1, 2)
        h = F.layer_norm(F.relu(h), (h.shape[-1],), sd[f"{p}.conv_layer.layer_norm_{n}.weight"],
                         sd[f"{p}.conv_layer.layer_norm_{n}.bias"], 1e-5)
    out = F.linear(h, sd[f"{p}.linear_layer.weight"], sd[f"{p}.linear_layer.bias"]).squeeze(-1)
    return out.masked_fill(pad, 0.0) if pad is not None else out


def bucket_index(values, bins):
    """torch.bucketize(v, bins, right=False): the number of boundaries strictly below v
    (model/modules.py:83,87,94,98)."""
    return torch.bucketize(values, bins)


def duration_rounded(log_d, d_control):
    """model/modules.py:132-135."""
    return torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0)


def repeat_counts(durations):
    """model/modules.py:186-187: max(int(d), 0) -- truncation toward zero."""
    return torch.clamp(torch.trunc(durations.double()).to(torch.int64), min=0)


def frame_to_phoneme_map(reps_row):
    """Index map of LengthRegulator.expand (model/modules.py:182-190) for one utterance, numpy int64."""
    reps_row = np.asarray(reps_row, dtype=np.int64)
    return np.repeat(np.arange(reps_row.shape[0], dtype=np.int64), reps_row)


def length_regulate(x, durations, max_len=None, loop=False):
    """model/modules.py:167-194 + utils/tools.py:360-378.  loop=True walks the rows one by
    one exactly as the reference does (used for the timed CPU baseline); loop=False uses
    repeat_interleave (same result, used by the tests)."""
    reps = repeat_counts(durations)
    outs = []
    for b in range(x.shape[0]):
        if loop:
            pieces = [x[b, j].expand(max(int(durations[b, j].item()), 0), -1) for j in range(x.shape[1])]
            outs.append(torch.cat(pieces, 0))
        else:
            outs.append(torch.repeat_interleave(x[b], reps[b], dim=0))
    mel_len = torch.tensor([o.shape[0] for o in outs], dtype=torch.int64)
    T = int(max_len) if max_len else int(mel_len.max().item())      # tools.py:361-364 (truthiness)
    out = torch.stack([F.pad(o, (0, 0, 0, T - o.shape[0])) for o in outs])
    return out, mel_len


def variance_adaptor(sd, x, src_pad, mel_pad, max_len, p_target, e_target, d_target,
                     p_control, e_control, d_control, taps=None, loop_lr=False,
                     pitch_level="phoneme_level", energy_level="phoneme_level"):
    """model/modules.py:102-158.  Note energy is scaled by p_control (modules.py:123-125 and
    :146-148): e_control is accepted and ignored, as in the reference.  `*_level` is
    preprocess.yaml's preprocessing.{pitch,energy}.feature (modules.py:28-35): phoneme_level
    features are predicted before the LengthRegulator (:114-125), frame_level ones after (:139-148)."""
    va = "variance_adaptor"
    pitch = energy = None

    def feature(name, x, target, pad):
        pred = variance_predictor(sd, f"{va}.{name}_predictor", x, pad)
        if target is not None:
            idx = bucket_index(target, sd[f"{va}.{name}_bins"])
        else:
            pred = pred * p_control
            idx = bucket_index(pred, sd[f"{va}.{name}_bins"])
        return pred, idx, x + F.embedding(idx, sd[f"{va}.{name}_embedding.weight"])

    log_d = variance_predictor(sd, f"{va}.duration_predictor", x, src_pad)
    p_idx = e_idx = None
    if pitch_level == "phoneme_level":
        pitch, p_idx, x = feature("pitch", x, p_target, src_pad)
    if energy_level == "phoneme_level":
        energy, e_idx, x = feature("energy", x, e_target, src_pad)
    if taps is not None:
        taps["va_x"], taps["p_idx"], taps["e_idx"] = x, p_idx, e_idx
    if d_target is not None:
        x, mel_len = length_regulate(x, d_target, max_len, loop=loop_lr)
        d_round = d_target
    else:
        d_round = duration_rounded(log_d, d_control)
        x, mel_len = length_regulate(x, d_round, max_len, loop=loop_lr)
        mel_pad = pad_mask(mel_len)
    if pitch_level == "frame_level":
        pitch, p_idx, x = feature("pitch", x, p_target, mel_pad)
    if energy_level == "frame_level":
        energy, e_idx, x = feature("energy", x, e_target, mel_pad)
    if taps is not None and "frame_level" in (pitch_level, energy_level):
        taps["va_frames"], taps["p_idx_f"], taps["e_idx_f"] = x, p_idx, e_idx
    return x, pitch, energy, log_d, d_round, mel_len, mel_pad


# ------------------------------------------------------------------ postnet + top level

def postnet(sd, mel):
    """transformer/Layers.py:129-137 -- 5 x (Conv1d k=5 pad=2 + BatchNorm1d eval), tanh on the first 4."""
    x = mel.transpose(1, 2)
    for j in range(N_POSTNET):
        c, bn = f"postnet.convolutions.{j}.0.conv", f"postnet.convolutions.{j}.1"
        w = sd[c + ".weight"]
        x = F.conv1d(x, w, sd[c + ".bias"], padding=(w.shape[-1] - 1) // 2)
        x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                         sd[bn + ".bias"], False, 0.1, 1e-5)
        if j < N_POSTNET - 1:
            x = torch.tanh(x)
    return x.transpose(1, 2)


@torch.no_grad()
def forward(sd, speakers, emotions, arousals, valences, texts, src_lens, max_src_len,
            mels=None, mel_lens=None, max_mel_len=None, p_targets=None, e_targets=None,
            d_targets=None, p_control=1.0, e_control=1.0, d_control=1.0, taps=None, loop_lr=False,
            pitch_level="phoneme_level", energy_level="phoneme_level"):
    """model/fastspeech2.py:73-149.  Returns the reference's 10-tuple."""
    src_pad = pad_mask(src_lens, max_src_len)
    mel_pad = pad_mask(mel_lens, max_mel_len) if mel_lens is not None else None
    x = encoder(sd, texts, src_pad, taps)
    spk, emo = conditioning(sd, speakers, emotions, arousals, valences)
    x = x + spk.unsqueeze(1).expand(-1, max_src_len, -1)
    x = x + emo.unsqueeze(1).expand(-1, max_src_len, -1)
    if taps is not None:
        taps["cond_x"] = x
    x, pitch, energy, log_d, d_round, mel_lens, mel_pad = variance_adaptor(
        sd, x, src_pad, mel_pad, max_mel_len, p_targets, e_targets, d_targets,
        p_control, e_control, d_control, taps, loop_lr, pitch_level, energy_level)
    if taps is not None:
        taps["lr_out"] = x
    x = decoder(sd, x, mel_pad, taps)
    mel = F.linear(x, sd["mel_linear.weight"], sd["mel_linear.bias"])
    post = postnet(sd, mel) + mel
    return mel, post, pitch, energy, log_d, d_round, src_pad, mel_pad, src_lens, mel_lens


def cast_state_dict(sd, dtype):
    """fp64 (ground truth) or fp32 copies of the floating tensors; integer buffers untouched."""
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
