"""`HiFiGANGeneratorB200` -- host-side mirror of the reference's HiFi-GAN `Generator`
(hifigan/models.py:112-174) plus `get_vocoder` / `vocoder_infer` (utils/model.py:37-92).

The module holds the generator's parameters under the reference's names (the remove_weight_norm() form; checkpoints
in weight_norm form -- `weight_g` / `weight_v`, as `generator_universal.pth.tar` stores them -- are folded on load) and
calls libfs2b200.so (`fs2_voc_*`, include/fs2_b200.h).  No PyTorch/CPU implementation exists here.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .synthetic import synthetic_vocoder_state_dict, vocoder_schema

HOP = 256


def fold_weight_norm(state_dict):
    """w = g * v / ||v|| (norm over every dim but 0) -- what Generator.remove_weight_norm() leaves (models.py:169-174)."""
    out = {}
    for k, v in state_dict.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            wv = state_dict[base + ".weight_v"]
            norm = wv.reshape(wv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base + ".weight"] = wv * (v / norm)
        elif not k.endswith(".weight_v"):
            out[k] = v
    return out


class HiFiGANGeneratorB200(nn.Module):
    """Drop-in for `hifigan.Generator(h)` in eval mode (V1 architecture of hifigan/config.json) on one B200."""

    def __init__(self, h=None, init_seed=0, math_mode="tf32"):
        super().__init__()
        self.math_mode = {"tf32": _lib.MATH_TF32, "bf16": _lib.MATH_BF16}[math_mode]
        if h is not None:
            want = dict(upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
                        resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3, resblock="1")
            for k, v in want.items():
                if k in h and h[k] != v:
                    raise ValueError(f"fs2_b200 vocoder kernels are specialised for hifigan/config.json ({k}={v}), got {h[k]}")
        init = synthetic_vocoder_state_dict(seed=init_seed)
        for key, _ in vocoder_schema():
            mod = self
            parts = key.split(".")
            for name in parts[:-1]:
                if name not in mod._modules:
                    mod.add_module(name, nn.Module())
                mod = mod._modules[name]
            mod.register_parameter(parts[-1], nn.Parameter(init[key], requires_grad=False))
        self._ctx = None
        self._ctx_device = None
        self._dirty = True
        self.eval()

    def load_state_dict(self, state_dict, strict=True, assign=False):
        if any(k.endswith(".weight_g") for k in state_dict):
            state_dict = fold_weight_norm(state_dict)
        res = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._dirty = True
        return res

    def remove_weight_norm(self):
        """Kept for call-site compatibility (utils/model.py:66): the weights are always held in plain form."""
        return self

    def _apply(self, fn, recurse=True):
        res = super()._apply(fn, recurse)
        self._dirty = True
        return res

    def train(self, mode=True):
        if mode:
            raise RuntimeError("HiFiGANGeneratorB200 is an inference engine")
        return super().train(False)

    def _device(self):
        return self.conv_post.weight.device

    def _ensure_ctx(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("HiFiGANGeneratorB200 runs on a CUDA device only: call .to('cuda') first (there is no CPU path)")
        lib = _lib.load_library()
        if self._ctx is not None and self._ctx_device != dev:
            lib.fs2_voc_destroy(self._ctx)
            self._ctx = None
        if self._ctx is None:
            ctx = C.c_void_p()
            code = lib.fs2_voc_create(dev.index if dev.index is not None else torch.cuda.current_device(), self.math_mode,
                                      C.byref(ctx))
            if code != 0:
                raise RuntimeError(f"libfs2b200 error {code}: {lib.fs2_voc_last_error(None).decode()}")
            self._ctx, self._ctx_device, self._dirty = ctx, dev, True
        if self._dirty:
            for key, t in self.state_dict().items():
                t = t.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                self._check(lib, lib.fs2_voc_set_weight(self._ctx, key.encode(), t.data_ptr(), shape, t.dim()))
            self._check(lib, lib.fs2_voc_prepare(self._ctx, torch.cuda.current_stream(dev).cuda_stream))
            self._dirty = False
        return lib

    def _check(self, lib, code):
        if code != 0:
            raise RuntimeError(f"libfs2b200 error {code}: {lib.fs2_voc_last_error(self._ctx).decode()}")

    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load_library().fs2_voc_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    @property
    def last_launch_count(self):
        return _lib.load_library().fs2_voc_last_launch_count(self._ctx) if self._ctx is not None else 0

    @torch.no_grad()
    def forward(self, x, mel_lens=None):
        """x: mel [B, 80, T] fp32 on the module's device (any strides: `postnet.transpose(1, 2)` is fine).
        Returns wav [B, 1, 256 T] like Generator.forward (models.py:148-167).  `mel_lens` (int64 [B]) is an extension:
        frames beyond it are skipped and the matching samples are 0 (see include/fs2_b200.h)."""
        lib = self._ensure_ctx()
        dev = self._device()
        if not torch.is_tensor(x) or x.device != dev or x.dim() != 3 or x.shape[1] != 80:
            raise RuntimeError(f"mel must be a [B, 80, T] tensor on {dev}")
        x = x.to(torch.float32)
        B, _, T = x.shape
        lens = None
        if mel_lens is not None:
            lens = mel_lens.to(dev, torch.int64).contiguous()
            if tuple(lens.shape) != (B,):
                raise RuntimeError("mel_lens must have shape [B]")
        wav = torch.empty(B, 1, T * HOP, dtype=torch.float32, device=dev)
        if T == 0:      # an all-zero-duration batch (the acoustic model returns [B, 0, 80]): nothing to synthesise
            return wav
        sb, sc, st = x.stride()
        self._check(lib, lib.fs2_voc_forward(self._ctx, torch.cuda.current_stream(dev).cuda_stream, x.data_ptr(), sb, sc, st,
                                             B, T, lens.data_ptr() if lens is not None else None, wav.data_ptr()))
        return wav


def get_vocoder(config, device, ckpt_path=None, random_init=False, math_mode="tf32"):
    """utils/model.py:37-71 (HiFi-GAN branch): build the generator, load `generator_<speaker>.pth.tar`["generator"]."""
    name = config["vocoder"]["model"]
    if name != "HiFi-GAN":
        raise ValueError("only the HiFi-GAN vocoder is implemented on the B200 engine")
    voc = HiFiGANGeneratorB200(math_mode=math_mode)
    if not random_init:
        speaker = config["vocoder"]["speaker"]
        path = ckpt_path or os.path.join("hifigan", f"generator_{speaker}.pth.tar")
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        voc.load_state_dict(ckpt["generator"])
    return voc.to(device).eval()


def vocoder_infer(mels, vocoder, model_config, preprocess_config, lengths=None, skip_padding=False):
    """utils/model.py:74-92: mels [B, 80, T] -> list of int16 numpy waveforms, trimmed to `lengths` samples.
    Default = the reference's semantics: the generator runs over the whole padded batch and the waveforms are trimmed
    afterwards, so the last receptive-field samples of a short utterance see the padded frames exactly as they do there.
    skip_padding=True is the ragged fast path (an extension): frames beyond `lengths` are not synthesised at all and
    every utterance is generated as if it were alone in the batch (zero padding after its last frame) -- about a third
    of the work at batch 64; it differs from the default only within the generator's receptive field of each tail."""
    mel_lens = None
    if lengths is not None:
        lengths = [int(n) for n in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        if skip_padding:
            mel_lens = torch.tensor([(n + HOP - 1) // HOP for n in lengths], dtype=torch.int64)
    wavs = vocoder(mels, mel_lens=mel_lens).squeeze(1)
    wavs = (wavs.cpu().numpy() * preprocess_config["preprocessing"]["audio"]["max_wav_value"]).astype("int16")
    wavs = [w for w in wavs]
    if lengths is not None:
        wavs = [w[:n] for w, n in zip(wavs, lengths)]
    return wavs
