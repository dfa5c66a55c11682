"""CPU emulation of the operand-rounding error budget (no GPU): which contractions must run as 3xTF32
for the `parity` mode to sit well below 1e-3 on the mel.  Operands of the selected ops are rounded to
TF32 (nearest, ties away; fp64 accumulation), everything else stays fp64; the frame side is teacher
forced exactly as in tests/test_gpu_forward.py.  Usage: python tools/error_budget.py [n_utts]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fs2_b200
from oracle import fs2_oracle as O


def tf32(x):
    f = x.to(torch.float32).contiguous()
    bits = f.view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32).to(x.dtype)


def split3(a, b, fn):
    """3xTF32: a = ah + al, b = bh + bl; ah*bh + al*bh + ah*bl"""
    ah, bh = tf32(a), tf32(b)
    al, bl = tf32(a - ah), tf32(b - bh)
    return fn(ah, bh) + fn(al, bh) + fn(ah, bl)


def run(mode_gemm, mode_attn, sd64, args, L, want):
    realF, realT = torch.nn.functional, torch

    def rnd(mode, a, b, fn):
        if mode == "exact":
            return fn(a, b)
        if mode == "tf32":
            return fn(tf32(a), tf32(b))
        return split3(a, b, fn)

    F = types.SimpleNamespace(**{k: getattr(realF, k) for k in dir(realF) if not k.startswith("__")})
    F.linear = lambda x, w, b=None: rnd(mode_gemm, x, w, lambda p, q: realF.linear(p, q)) + (0 if b is None else b)
    F.conv1d = lambda x, w, b=None, padding=0: rnd(mode_gemm, x, w, lambda p, q: realF.conv1d(p, q, None, padding=padding)) + \
        (0 if b is None else b.view(1, -1, 1))
    T = types.SimpleNamespace(**{k: getattr(realT, k) for k in dir(realT) if not k.startswith("__")})
    T.bmm = lambda a, b: rnd(mode_attn, a, b, realT.bmm)
    O.F, O.torch = F, T
    try:
        got = O.forward(sd64, *args, L, d_targets=want[5], p_targets=want[2], e_targets=want[3], mel_lens=want[9],
                        max_mel_len=int(want[9].max()))
    finally:
        O.F, O.torch = realF, realT
    errs = []
    for i in (0, 1):
        errs.append(max(float((got[i][b, :t] - want[i][b, :t]).abs().max()) for b, t in enumerate(want[9].tolist())))
    return errs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    syn = fs2_b200.synthetic
    sd64 = O.cast_state_dict(syn.synthetic_state_dict(seed=0), torch.float64)
    batch = syn.config2_batch(seed=0)
    args = [batch[k][:n] for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
    L = int(args[5].max())
    args[4] = args[4][:, :L].contiguous()
    with torch.no_grad():
        want = O.forward(sd64, *args, L)
        for g, a in (("tf32", "tf32"), ("3x", "tf32"), ("tf32", "exact"), ("3x", "3x"), ("exact", "tf32")):
            mel, post = run(g, a, sd64, args, L, want)
            print(f"gemm={g:5s} attention={a:5s}  mel max-abs {mel:.2e}  postnet {post:.2e}", flush=True)


if __name__ == "__main__":
    main()
