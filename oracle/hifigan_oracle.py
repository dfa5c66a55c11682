"""CPU ORACLE for the HiFi-GAN generator  --  TEST INFRASTRUCTURE ONLY (see oracle/fs2_oracle.py for the rules:
imported by tests/, smoke() and bench.py's CPU legs only, never by the product package).

Flat-function restatement of `Generator.forward` (/root/reference/hifigan/models.py:148-167) and `ResBlock.forward`
(:96-103) over a state dict in the remove_weight_norm() form (plain `.weight` / `.bias`), architecture from
hifigan/config.json (V1: upsample 8,8,2,2 with kernels 16,16,4,4 from 512 channels; ResBlocks k = 3,7,11 with
dilations 1,3,5).  Same torch CPU operators as the reference (F.conv1d, F.conv_transpose1d, F.leaky_relu, tanh).

Pinning: the reference has no test or golden vector for the vocoder and ships no vocoder weights
(.MISSING_LARGE_BLOBS); this file is pinned against the output of the unmodified reference `hifigan.Generator`
run in float64 on seeded weights (tests/golden/make_golden_vocoder.py -> tests/golden/vocoder.npz), <= 1e-12.
"""
import torch
import torch.nn.functional as F

UP_RATES = (8, 8, 2, 2)          # hifigan/config.json:11
UP_KERNELS = (16, 16, 4, 4)      # config.json:12
RES_KERNELS = (3, 7, 11)         # config.json:14
RES_DILATIONS = (1, 3, 5)        # config.json:15
LRELU_SLOPE = 0.1                # models.py:7


def fold_weight_norm(sd):
    """weight_norm checkpoints (`weight_g`, `weight_v`; models.py:24-90,117-135) -> plain weights, as
    Generator.remove_weight_norm() does (models.py:169-174): w = g * v / ||v|| with the norm over all dims but 0."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            wv = sd[base + ".weight_v"]
            norm = wv.reshape(wv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base + ".weight"] = wv * (v / norm)
        elif not k.endswith(".weight_v"):
            out[k] = v
    return out


def resblock(sd, p, x, k):
    """models.py:96-103."""
    for m, d in enumerate(RES_DILATIONS):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs1.{m}.weight"], sd[f"{p}.convs1.{m}.bias"], dilation=d, padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{p}.convs2.{m}.weight"], sd[f"{p}.convs2.{m}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


@torch.no_grad()
def generator(sd, mel):
    """models.py:148-167.  mel [B, 80, T] -> wav [B, 1, 256 T]."""
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(UP_RATES, UP_KERNELS)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RES_KERNELS):
            r = resblock(sd, f"resblocks.{i * len(RES_KERNELS) + j}", x, rk)
            xs = r if xs is None else xs + r
        x = xs / len(RES_KERNELS)
    x = F.leaky_relu(x)                       # default slope 0.01 (models.py:163)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def vocoder_infer(sd, mels, max_wav_value=32768.0, lengths=None):
    """utils/model.py:74-92 (HiFi-GAN branch): int16 waveforms, optionally trimmed to `lengths` samples."""
    wavs = (generator(sd, mels).squeeze(1).cpu().numpy() * max_wav_value).astype("int16")
    wavs = [w for w in wavs]
    if lengths is not None:
        wavs = [w[: int(n)] for w, n in zip(wavs, lengths)]
    return wavs
