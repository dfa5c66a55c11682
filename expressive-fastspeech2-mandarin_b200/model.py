"""`FastSpeech2B200` -- the host-side mirror of the reference's `FastSpeech2` module.

Same constructor arguments, same 240 state-dict keys and shapes, same
`forward` signature and 10-tuple result as model/fastspeech2.py:16-149, so
`model.load_state_dict(ckpt["model"])` and `model(*(batch[2:]), p_control=..., ...)`
(synthesize_chinese_pinyin.py:138-145) work unchanged.  The module only holds the
parameters; every computation happens in libfs2b200.so (include/fs2_b200.h).  There is no
PyTorch/CPU implementation of the forward here and none is ever selected: without the
library, or with CPU tensors, the calls raise.
"""
import ctypes as C
import json
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .config import N_SRC_VOCAB, POSTNET_DIM, POSTNET_KERNEL, POSTNET_LAYERS, check_supported
from .synthetic import sinusoid_table

_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


def _schema(n_speaker, n_emotion, n_arousal, n_valence, max_seq_len):
    """(key, shape) of every tensor of the reference's state dict, in its order."""
    d, inner = 256, 1024
    out = [("encoder.position_enc", (1, max_seq_len + 1, d)), ("encoder.src_word_emb.weight", (N_SRC_VOCAB, d))]

    def fft(prefix):
        for w in ("w_qs", "w_ks", "w_vs"):
            out.extend([(f"{prefix}.slf_attn.{w}.weight", (d, d)), (f"{prefix}.slf_attn.{w}.bias", (d,))])
        out.extend([(f"{prefix}.slf_attn.layer_norm.weight", (d,)), (f"{prefix}.slf_attn.layer_norm.bias", (d,)),
                    (f"{prefix}.slf_attn.fc.weight", (d, d)), (f"{prefix}.slf_attn.fc.bias", (d,)),
                    (f"{prefix}.pos_ffn.w_1.weight", (inner, d, 9)), (f"{prefix}.pos_ffn.w_1.bias", (inner,)),
                    (f"{prefix}.pos_ffn.w_2.weight", (d, inner, 1)), (f"{prefix}.pos_ffn.w_2.bias", (d,)),
                    (f"{prefix}.pos_ffn.layer_norm.weight", (d,)), (f"{prefix}.pos_ffn.layer_norm.bias", (d,))])

    for i in range(4):
        fft(f"encoder.layer_stack.{i}")
    va = "variance_adaptor"
    out.extend([(f"{va}.pitch_bins", (255,)), (f"{va}.energy_bins", (255,))])
    for name in ("duration", "pitch", "energy"):
        p = f"{va}.{name}_predictor"
        for n in (1, 2):
            out.extend([(f"{p}.conv_layer.conv1d_{n}.conv.weight", (d, d, 3)), (f"{p}.conv_layer.conv1d_{n}.conv.bias", (d,)),
                        (f"{p}.conv_layer.layer_norm_{n}.weight", (d,)), (f"{p}.conv_layer.layer_norm_{n}.bias", (d,))])
        out.extend([(f"{p}.linear_layer.weight", (1, d)), (f"{p}.linear_layer.bias", (1,))])
    out.extend([(f"{va}.pitch_embedding.weight", (256, d)), (f"{va}.energy_embedding.weight", (256, d))])
    out.append(("decoder.position_enc", (1, max_seq_len + 1, d)))
    for i in range(6):
        fft(f"decoder.layer_stack.{i}")
    out.extend([("mel_linear.weight", (80, d)), ("mel_linear.bias", (80,))])
    chans = [80] + [POSTNET_DIM] * (POSTNET_LAYERS - 1) + [80]
    for j in range(POSTNET_LAYERS):
        c = f"postnet.convolutions.{j}"
        out.extend([(f"{c}.0.conv.weight", (chans[j + 1], chans[j], POSTNET_KERNEL)), (f"{c}.0.conv.bias", (chans[j + 1],))])
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{c}.1.{leaf}", (chans[j + 1],)))
        out.append((f"{c}.1.num_batches_tracked", ()))
    out.extend([("speaker_emb.weight", (n_speaker, d)), ("emotion_emb.weight", (n_emotion, d // 2)),
                ("arousal_emb.weight", (n_arousal, d // 4)), ("valence_emb.weight", (n_valence, d // 4)),
                ("emotion_linear.0.weight", (d, d)), ("emotion_linear.0.bias", (d,))])
    return out


class FastSpeech2B200(nn.Module):
    """Drop-in for `FastSpeech2(preprocess_config, model_config)` in eval mode on one B200."""

    def __init__(self, preprocess_config, model_config, math_mode="tf32", init_seed=0):
        super().__init__()
        check_supported(preprocess_config, model_config)
        self.model_config = model_config
        root = preprocess_config["path"]["preprocessed_path"]
        with open(os.path.join(root, "speakers.json")) as f:       # model/fastspeech2.py:31-37
            n_speaker = len(json.load(f))
        with open(os.path.join(root, "emotions.json")) as f:       # model/fastspeech2.py:45-54
            emo = json.load(f)
        with open(os.path.join(root, "stats.json")) as f:          # model/modules.py:41-46
            stats = json.load(f)
        self._dims = dict(n_src_vocab=N_SRC_VOCAB, n_speaker=n_speaker, n_emotion=len(emo["emotion_dict"]),
                          n_arousal=len(emo["arousal_dict"]), n_valence=len(emo["valence_dict"]),
                          max_seq_len=int(model_config["max_seq_len"]))
        pp = preprocess_config["preprocessing"]                    # model/modules.py:28-35
        self.pitch_frame_level = pp["pitch"]["feature"] == "frame_level"
        self.energy_frame_level = pp["energy"]["feature"] == "frame_level"
        # "tf32" (default), "bf16" (bf16 operands) or "parity" (split-operand 3xTF32 contractions, ~1e-4 of fp64)
        self.math_mode = {"tf32": _lib.MATH_TF32, "bf16": _lib.MATH_BF16, "parity": _lib.MATH_TF32X3,
                          "tf32x3": _lib.MATH_TF32X3}[math_mode]

        # Parameter tree built from the schema; values: deterministic tables where the reference
        # computes them (position_enc, bins), seeded random elsewhere (the reference random-inits).
        from .synthetic import synthetic_state_dict
        init = synthetic_state_dict(seed=init_seed, duration_bias=0.0)
        ve = model_config["variance_embedding"]

        def bins(lo, hi, kind):                                     # model/modules.py:48-71
            if kind == "log":
                return torch.exp(torch.linspace(np.log(lo), np.log(hi), ve["n_bins"] - 1))
            return torch.linspace(lo, hi, ve["n_bins"] - 1)

        init["variance_adaptor.pitch_bins"] = bins(stats["pitch"][0], stats["pitch"][1], ve["pitch_quantization"])
        init["variance_adaptor.energy_bins"] = bins(stats["energy"][0], stats["energy"][1], ve["energy_quantization"])
        pe = sinusoid_table(self._dims["max_seq_len"] + 1, 256).unsqueeze(0)
        for key, shape in _schema(self._dims["n_speaker"], self._dims["n_emotion"], self._dims["n_arousal"],
                                  self._dims["n_valence"], self._dims["max_seq_len"]):
            if key.endswith("position_enc"):
                value = pe.clone()
            elif key in init and tuple(init[key].shape) == tuple(shape):
                value = init[key]
            else:  # a table whose size differs from the synthetic fixture
                value = torch.randn(shape, generator=torch.Generator().manual_seed(init_seed)) if shape else torch.tensor(0)
            self._attach(key, value)

        self._ctx = None
        self._ctx_device = None
        self._dirty = True
        self._pinned = {}
        self.eval()

    # ------------------------------------------------------------------ parameter tree
    def _attach(self, key, value):
        parts = key.split(".")
        mod = self
        for name in parts[:-1]:
            if name not in mod._modules:
                mod.add_module(name, nn.Module())
            mod = mod._modules[name]
        if parts[-1] in _BUFFER_SUFFIXES:
            mod.register_buffer(parts[-1], value)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))

    def load_state_dict(self, state_dict, strict=True, assign=False):
        res = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._dirty = True
        return res

    def _apply(self, fn, recurse=True):
        res = super()._apply(fn, recurse)
        self._dirty = True
        return res

    def train(self, mode=True):
        if mode:
            raise RuntimeError("FastSpeech2B200 is an inference engine: training mode (dropout, BatchNorm updates, "
                               "backward) is outside the accelerated path")
        return super().train(False)

    def refresh_weights(self):
        """Re-upload the parameters to the library (call after modifying them in place)."""
        self._dirty = True

    # ------------------------------------------------------------------ library context
    def _device(self):
        return self.mel_linear.weight.device

    def _ensure_ctx(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("FastSpeech2B200 runs on a CUDA device only: call .to('cuda') first (there is no CPU path)")
        lib = _lib.load_library()
        if self._ctx is not None and self._ctx_device != dev:
            lib.fs2_destroy(self._ctx)
            self._ctx = None
        if self._ctx is None:
            cfg = _lib.Config(math_mode=self.math_mode, pitch_frame_level=int(self.pitch_frame_level),
                              energy_frame_level=int(self.energy_frame_level), **self._dims)
            ctx = C.c_void_p()
            code = lib.fs2_create(C.byref(cfg), dev.index if dev.index is not None else torch.cuda.current_device(),
                                  C.byref(ctx))
            _lib.check(lib, None, code)
            self._ctx, self._ctx_device, self._dirty = ctx, dev, True
        if self._dirty:
            stream = torch.cuda.current_stream(dev).cuda_stream
            for key, t in self.state_dict().items():
                if not t.is_floating_point():
                    continue
                t = t.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(lib, self._ctx, lib.fs2_set_weight(self._ctx, key.encode(), t.data_ptr(), shape, t.dim()))
            _lib.check(lib, self._ctx, lib.fs2_prepare(self._ctx, stream))
            self._dirty = False
        return lib

    def __del__(self):
        try:
            if self._ctx is not None:
                _lib.load_library().fs2_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def debug_taps(self, on=True):
        lib = self._ensure_ctx()
        lib.fs2_debug_enable(self._ctx, 1 if on else 0)

    def fetch_tap(self, name):
        lib = self._ensure_ctx()
        rows, cols = C.c_int64(), C.c_int64()
        _lib.check(lib, self._ctx, lib.fs2_debug_fetch(self._ctx, name.encode(), None, 0, C.byref(rows), C.byref(cols)))
        dtype = np.int32 if name.endswith("_start") else np.float32
        out = np.empty((rows.value, cols.value), dtype=dtype)
        _lib.check(lib, self._ctx, lib.fs2_debug_fetch(self._ctx, name.encode(), out.ctypes.data, out.nbytes, None, None))
        return out

    @property
    def last_launch_count(self):
        return _lib.load_library().fs2_last_launch_count(self._ctx) if self._ctx is not None else 0

    # ------------------------------------------------------------------ forward
    def _idx(self, t, name, shape):
        if not torch.is_tensor(t) or t.device != self._device():
            raise RuntimeError(f"{name} must be a tensor on {self._device()} (got {getattr(t, 'device', type(t))})")
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t.to(torch.int64).contiguous()

    def _target(self, t, name, shape):
        if t is None:
            return None
        if not torch.is_tensor(t) or t.device != self._device():
            raise RuntimeError(f"{name} must be a tensor on {self._device()}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        return t.to(torch.float32).contiguous()

    def _stage1(self, speakers, emotions, arousals, valences, texts, src_lens, max_src_len, max_mel_len, p_targets, e_targets,
                d_targets, p_control, e_control, d_control, eager=True):
        """fs2_forward_stage1: masks, encoder, conditioning, variance adaptor up to the durations.  Returns the phoneme-side
        tensors and the frame-side sizes (the call blocks for them)."""
        if self.training:
            raise RuntimeError("FastSpeech2B200.forward is inference-only")
        lib = self._ensure_ctx()
        dev = self._device()
        B, L = int(texts.shape[0]), int(max_src_len)
        if texts.dim() != 2 or int(texts.shape[1]) != L:
            # the reference sizes the encoder from texts.shape[1] and the masks from max_src_len and
            # fails on a mismatch
            raise RuntimeError(f"texts.shape[1] ({tuple(texts.shape)}) must equal max_src_len ({L})")
        spk = self._idx(speakers, "speakers", (B,))
        emo = self._idx(emotions, "emotions", (B,))
        aro = self._idx(arousals, "arousals", (B,))
        val = self._idx(valences, "valences", (B,))
        txt = self._idx(texts, "texts", (B, L))
        lens = self._idx(src_lens, "src_lens", (B,))
        # frame_level targets live on the frame axis ([B, max_mel_len]); their shape is checked once T is known
        p_t = self._target(p_targets, "p_targets", None if self.pitch_frame_level else (B, L))
        e_t = self._target(e_targets, "e_targets", None if self.energy_frame_level else (B, L))
        d_t = self._target(d_targets, "d_targets", (B, L))

        f32 = dict(dtype=torch.float32, device=dev)
        # one allocation for the four [B, L] prediction tensors: the host-buffer entry reads them back with one copy
        pred = torch.empty(4, B, L, **f32)
        pitch, energy, log_d, d_round = pred[0], pred[1], pred[2], pred[3]
        self._last_pred = pred
        src_mask = torch.empty(B, L, dtype=torch.bool, device=dev)
        out_lens = torch.empty(B, dtype=torch.int64, device=dev)

        def ptr(t):
            return t.data_ptr() if t is not None else None

        inp = _lib.Inputs(batch=B, max_src_len=L, speakers=ptr(spk), emotions=ptr(emo), arousals=ptr(aro),
                          valences=ptr(val), texts=ptr(txt), src_lens=ptr(lens), p_targets=ptr(p_t), e_targets=ptr(e_t),
                          d_targets=ptr(d_t), p_control=float(p_control), e_control=float(e_control),
                          d_control=float(d_control), max_mel_len=int(max_mel_len) if max_mel_len else 0)
        s1 = _lib.Stage1Out(pitch=ptr(pitch), energy=ptr(energy), log_d=ptr(log_d), d_rounded=ptr(d_round),
                            src_mask=ptr(src_mask), mel_lens=ptr(out_lens))
        stream = torch.cuda.current_stream(dev).cuda_stream
        # the targets must outlive stage 2 (frame_level features and the fused energy add read them there)
        self._keep = (spk, emo, aro, val, txt, lens, p_t, e_t, d_t)
        # stage 2 up to the PostNet is enqueued by stage 1 itself (the device does not wait for the output allocations
        # below).  An asynchronous host read of the previous call's packed rows (synthesize_host_async), which that body
        # overwrites, is waited for on the device right before the body: the library gets its event.
        pending = getattr(self, "_pending_read", None)
        eager = eager and getattr(self, "eager_stage2", True)      # (attribute: tests switch the mode off to compare)
        lib.fs2_set_eager_stage2(self._ctx, 1 if eager else 0)
        if pending is not None and pending.cuda_event:
            # (the library enqueues the wait at the end of stage 1, whether or not the body follows)
            lib.fs2_set_stage2_wait_event(self._ctx, C.c_void_p(pending.cuda_event))
            self._pending_read = None        # consumed by the library; `pending` keeps the event alive through the call
        elif pending is not None:
            lib.fs2_set_eager_stage2(self._ctx, 0)      # the wait stays where it was: before stage 2, in _stage2
        _lib.check(lib, self._ctx, lib.fs2_forward_stage1(self._ctx, stream, C.byref(inp), C.byref(s1)))
        self.last_total_frames = int(s1.total_frames)
        return dict(B=B, L=L, T=int(s1.max_mel_len), pitch=pitch, energy=energy, log_d=log_d, d_round=d_round,
                    src_mask=src_mask, mel_lens=out_lens, p_t=p_t, e_t=e_t)

    def _stage2(self, B, T, pitch=None, energy=None, p_t=None, e_t=None):
        """fs2_forward_stage2: length regulator, decoder, mel_linear, PostNet into freshly allocated outputs."""
        lib = self._ensure_ctx()
        dev = self._device()
        f32 = dict(dtype=torch.float32, device=dev)
        mel = torch.empty(B, T, 80, **f32)
        post = torch.empty(B, T, 80, **f32)
        mel_mask = torch.empty(B, T, dtype=torch.bool, device=dev)
        for flag, t, name in ((self.pitch_frame_level, p_t, "p_targets"), (self.energy_frame_level, e_t, "e_targets")):
            if flag and t is not None and tuple(t.shape) != (B, T):
                raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {(B, T)} (frame_level feature)")
        if self.pitch_frame_level:                                  # model/modules.py:139-143: prediction is [B, T]
            pitch = torch.empty(B, T, **f32)
        if self.energy_frame_level:                                 # model/modules.py:144-148
            energy = torch.empty(B, T, **f32)
        pending = getattr(self, "_pending_read", None)
        if pending is not None:       # an asynchronous host read of the previous call's packed rows (synthesize_host_async)
            torch.cuda.current_stream(dev).wait_event(pending)
            self._pending_read = None
        io = _lib.Stage2IO(mel=mel.data_ptr(), postnet=post.data_ptr(), mel_mask=mel_mask.data_ptr(),
                           pitch_frames=pitch.data_ptr() if self.pitch_frame_level else None,
                           energy_frames=energy.data_ptr() if self.energy_frame_level else None)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib, self._ctx, lib.fs2_forward_stage2(self._ctx, stream, C.byref(io)))
        return mel, post, mel_mask, pitch, energy

    @torch.no_grad()
    def forward(self, speakers, emotions, arousals, valences, texts, src_lens, max_src_len, mels=None, mel_lens=None,
                max_mel_len=None, p_targets=None, e_targets=None, d_targets=None, p_control=1.0, e_control=1.0,
                d_control=1.0):
        s1 = self._stage1(speakers, emotions, arousals, valences, texts, src_lens, max_src_len, max_mel_len, p_targets,
                          e_targets, d_targets, p_control, e_control, d_control)
        if d_targets is not None and mel_lens is not None:
            if not torch.equal(mel_lens.to(s1["mel_lens"].device, torch.int64), s1["mel_lens"]):
                raise RuntimeError("mel_lens must equal the row sums of trunc(d_targets) (inconsistent teacher forcing)")
        mel, post, mel_mask, pitch, energy = self._stage2(s1["B"], s1["T"], s1["pitch"], s1["energy"], s1["p_t"], s1["e_t"])
        return (mel, post, pitch, energy, s1["log_d"], d_targets if d_targets is not None else s1["d_round"], s1["src_mask"],
                mel_mask, src_lens, s1["mel_lens"])

    # ------------------------------------------------------------------ sharded batches re-balanced by frames (partition.py)
    @torch.no_grad()
    def encode(self, speakers, emotions, arousals, valences, texts, src_lens, max_src_len, p_control=1.0, e_control=1.0,
               d_control=1.0):
        """Stage 1 only (model/fastspeech2.py:92-131 up to the durations).  Returns a dict with the phoneme-side outputs
        (`pitch`, `energy`, `log_d`, `d_round`, `src_mask`), `mel_lens` [B] (device) and what the length regulator consumes,
        exported for another context: `hidden` [B, L, 256] (x + pitch embedding + energy embedding) and `reps` [B, L] int32."""
        s1 = self._stage1(speakers, emotions, arousals, valences, texts, src_lens, max_src_len, None, None, None, None,
                          p_control, e_control, d_control, eager=False)   # (stage 2 runs elsewhere, after the exchange)
        dev = self._device()
        hidden = torch.empty(s1["B"], s1["L"], 256, dtype=torch.float32, device=dev)
        reps = torch.empty(s1["B"], s1["L"], dtype=torch.int32, device=dev)
        lib = _lib.load_library()
        _lib.check(lib, self._ctx, lib.fs2_export_stage1(self._ctx, torch.cuda.current_stream(dev).cuda_stream,
                                                          hidden.data_ptr(), reps.data_ptr()))
        s1["hidden"], s1["reps"] = hidden, reps
        return s1

    @torch.no_grad()
    def decode(self, hidden, reps, src_lens, max_mel_len=None):
        """Stage 2 from exported rows (model/modules.py:136-137, fastspeech2.py:133-136) -- possibly of utterances another
        GPU encoded.  hidden [B, L, 256] fp32, reps [B, L] int32, src_lens [B] int64, all on this module's device.
        Returns (mel, postnet, mel_mask, mel_lens)."""
        lib = self._ensure_ctx()
        dev = self._device()
        B, L = int(hidden.shape[0]), int(hidden.shape[1])
        if tuple(hidden.shape) != (B, L, 256) or tuple(reps.shape) != (B, L) or hidden.device != dev or reps.device != dev:
            raise RuntimeError(f"decode: hidden [B, L, 256] and reps [B, L] must live on {dev}")
        hidden = hidden.to(torch.float32).contiguous()
        reps = reps.to(torch.int32).contiguous()
        lens = self._idx(src_lens, "src_lens", (B,))
        mel_lens = torch.empty(B, dtype=torch.int64, device=dev)
        total, t_max = C.c_int64(), C.c_int32()
        self._keep = (hidden, reps, lens)
        _lib.check(lib, self._ctx, lib.fs2_import_stage1(self._ctx, torch.cuda.current_stream(dev).cuda_stream, hidden.data_ptr(),
                                                          reps.data_ptr(), lens.data_ptr(), B, L, int(max_mel_len) if max_mel_len else 0,
                                                          mel_lens.data_ptr(), C.byref(total), C.byref(t_max)))
        self.last_total_frames = int(total.value)
        mel, post, mel_mask, _, _ = self._stage2(B, int(t_max.value))
        return mel, post, mel_mask, mel_lens

    # ------------------------------------------------------------------ host-buffer entry (what the CLI does)
    def _pinned_buf(self, role, dtype, numel):
        """One grow-only pinned staging buffer per ROLE (never per shape): a server with dynamic batch sizes keeps a
        handful of buffers sized for the largest request seen, instead of one buffer per distinct shape."""
        buf = self._pinned.get(role)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            cap = max(int(numel), int(buf.numel() * 1.5) if buf is not None else 0, 16)
            buf = torch.empty(cap, dtype=dtype).pin_memory()
            self._pinned[role] = buf
        return buf[:numel]

    def synthesize_host(self, batch, p_control=1.0, e_control=1.0, d_control=1.0, padded=False, copy=True):
        """`to_device(batch)` (utils/tools.py:117-127) + forward + the device->host reads that `synth_samples`
        performs (utils/tools.py:228-243: predictions[1] mel rows, [2] pitch, [3] energy, [5] durations, [9] mel_lens),
        on pinned staging buffers.  `batch` holds host numpy arrays: speakers, emotions, arousals, valences, texts,
        src_lens, max_src_len.

        Returns (mels, mel_lens, h2d_bytes, d2h_bytes): `mels` is a list of per-utterance [mel_len, 80] arrays (the PACKED
        rows `synth_samples` slices out of the padded tensor; a third of its bytes at batch 64), or with padded=True the
        padded [B, T, 80] array itself.  The pitch / energy / duration predictions of the same call are in
        `self.last_host` (dict of numpy arrays `pitch`, `energy`, `log_d`, `durations`).

        copy=True (default): every returned array owns its memory.  copy=False returns VIEWS into the pinned staging
        buffers, which a later call overwrites (the call after next: the staging is double buffered) -- only for callers
        that consume the result before calling again (the benchmark does).

        `synthesize_host_async` is the same call without the final wait: it returns a handle whose `.wait()` gives this
        tuple, so a serving loop can submit batch i+1 while the device->host copies of batch i are still in flight."""
        return self.synthesize_host_async(batch, p_control, e_control, d_control, padded=padded, copy=copy).wait()

    def synthesize_host_async(self, batch, p_control=1.0, e_control=1.0, d_control=1.0, padded=False, copy=True):
        """Enqueue one host-buffer synthesis and return a `HostResult` handle.  The device->host reads run on a side
        stream behind an event, into one of two staging slots, so they overlap the next call's encoder; the next call's
        stage 2 (which overwrites the library's frame-side buffer the packed rows are read from) waits for them."""
        dev = self._device()
        names = ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")
        # one pinned staging buffer and ONE host->device copy for the six int64 inputs; the device tensors are views of it
        arrays = [np.ascontiguousarray(batch[n], dtype=np.int64) for n in names]
        total = sum(a.size for a in arrays)
        stage = self._pinned_buf("inputs", torch.int64, total)
        stage_np, off = stage.numpy(), 0
        for a in arrays:
            stage_np[off: off + a.size] = a.reshape(-1)
            off += a.size
        on_dev = stage.to(dev, non_blocking=True)     # (consumed before stage 1 returns: it ends with a stream sync)
        h2d = total * 8
        dev_t, off = {}, 0
        for n, a in zip(names, arrays):
            dev_t[n] = on_dev[off: off + a.size].view(a.shape)
            off += a.size
        out = self.forward(dev_t["speakers"], dev_t["emotions"], dev_t["arousals"], dev_t["valences"], dev_t["texts"],
                           dev_t["src_lens"], int(batch["max_src_len"]), p_control=p_control, e_control=e_control,
                           d_control=d_control)
        post, lens = out[1], out[9]
        B, L = int(post.shape[0]), int(batch["max_src_len"])
        self.last_postnet = post      # device tensor [B, T, 80]: what the vocoder consumes next (utils/tools.py:258-262)
        compute = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(device=dev)
        side = self._copy_stream
        slot = self._host_slot = 1 - getattr(self, "_host_slot", 1)
        tag = lambda role: f"{role}#{slot}"
        side.wait_stream(compute)
        res = HostResult(self, copy, padded, h2d, B)
        res.keep = (out, self._last_pred)           # the device tensors outlive the copies
        with torch.cuda.stream(side):
            res.hl = self._pinned_buf(tag("lens"), torch.int64, B)
            res.hl.copy_(lens, non_blocking=True)
            # pitch / energy / log-duration / rounded duration: one copy of the [4, B, L] block the forward allocated
            # (frame_level features live on the frame axis and are read separately)
            res.hpred = self._pinned_buf(tag("pred"), torch.float32, 4 * B * L).view(4, B, L)
            res.hpred.copy_(self._last_pred, non_blocking=True)
            res.d2h = 4 * B * L * 4 + B * 8
            for key, idx, flag in (("pitch", 2, self.pitch_frame_level), ("energy", 3, self.energy_frame_level)):
                if flag:
                    t = out[idx]
                    hb = self._pinned_buf(tag("frame_" + key), torch.float32, t.numel()).view(t.shape)
                    hb.copy_(t, non_blocking=True)
                    res.frame_feats[key] = hb
                    res.d2h += t.numel() * 4
            if padded:
                res.hp = self._pinned_buf(tag("post"), torch.float32, post.numel()).view(post.shape)
                res.hp.copy_(post, non_blocking=True)
                res.d2h += post.numel() * 4
            else:
                # packed rows straight out of the library's frame-side buffer: rows = sum(mel_lens) + 12 reserved per utterance
                rows_cap = int(self.last_total_frames) + 12 * (B + 1)
                res.hp = self._pinned_buf(tag("packed"), torch.float32, rows_cap * 80).view(rows_cap, 80)
                res.hs = self._pinned_buf(tag("starts"), torch.int32, B + 1)
                rows = C.c_int64()
                lib = _lib.load_library()
                _lib.check(lib, self._ctx, lib.fs2_read_packed_postnet(self._ctx, side.cuda_stream, res.hp.data_ptr(), rows_cap,
                                                                        res.hs.data_ptr(), C.byref(rows)))
                res.d2h += int(rows.value) * 80 * 4 + (B + 1) * 4
            res.done = torch.cuda.Event(enable_timing=True)
            res.done.record(side)
        self._pending_read = res.done          # the next stage 2 must not overwrite the frame-side buffer before this
        return res


class HostResult:
    """Handle of one `synthesize_host_async` call; `wait()` blocks until the device->host copies have landed and returns
    (mels, mel_lens, h2d_bytes, d2h_bytes) as `synthesize_host` does (and fills `model.last_host`)."""

    def __init__(self, model, copy, padded, h2d, batch):
        self.model, self.copy, self.padded, self.h2d, self.B = model, copy, padded, h2d, batch
        self.frame_feats, self.hs, self.d2h, self.done, self.keep = {}, None, 0, None, None

    def wait(self):
        self.done.synchronize()
        own = (lambda a: np.array(a)) if self.copy else (lambda a: a)
        lens_np = self.hl.numpy()
        if self.padded:
            mels = own(self.hp.numpy())
        else:
            packed, starts = self.hp.numpy(), self.hs.numpy()
            mels = [own(packed[int(starts[b]): int(starts[b]) + int(lens_np[b])]) for b in range(self.B)]
        pred_np = self.hpred.numpy()
        ff = self.frame_feats
        self.model.last_host = {"pitch": own(ff["pitch"].numpy() if "pitch" in ff else pred_np[0]),
                                "energy": own(ff["energy"].numpy() if "energy" in ff else pred_np[1]),
                                "log_d": own(pred_np[2]), "durations": own(pred_np[3])}
        self.keep = None
        return mels, own(lens_np), self.h2d, self.d2h


def get_model(preprocess_config, model_config, state_dict=None, device="cuda", **kw):
    """utils/model.py:11-34 without the file I/O: construct, load, move, eval."""
    model = FastSpeech2B200(preprocess_config, model_config, **kw)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model.to(device).eval()
