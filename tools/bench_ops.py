"""Time the GEMM shapes of the forward in isolation through the C ABI (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream

ROWS = int(os.environ.get("ROWS", 26788))
SHAPES = [  # name, K, N, taps, act, residual, ln
    ("qkv", 256, 768, 1, 0, False, False),
    ("fc_ln", 256, 256, 1, 0, True, True),
    ("conv9", 256, 1024, 9, 1, False, False),
    ("w2_ln", 1024, 256, 1, 0, True, True),
    ("pn_512_512", 512, 512, 5, 2, False, False),
    ("pn_80_512", 80, 512, 5, 2, False, False),
    ("pn_512_80", 512, 80, 5, 0, True, False),
    ("mel", 256, 80, 1, 0, False, False),
]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
L = lib()
for engines in [(1,)]:
    for name, K, N, taps, act, use_res, ln in SHAPES:
        g = torch.Generator().manual_seed(1)
        A = torch.randn(ROWS, K, generator=g).to(DEV)
        W = (torch.randn(taps, N, K, generator=g) / np.sqrt(K * taps)).to(DEV)
        bias = torch.randn(N, generator=g).to(DEV)
        res = torch.randn(ROWS, N, generator=g).to(DEV) if use_res else None
        gamma, beta = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
        out = torch.empty(ROWS, N, device=DEV)
        line = f"{name:12s} rows={ROWS} K={K} N={N} taps={taps}"
        for eng in engines:
            if ln and eng != 1:
                continue
            def call():
                if ln:
                    return L.fs2_op_conv_gemm_ln(stream(), ptr(A), K, ROWS, ptr(W), ptr(bias), taps, (taps - 1) // 2, K, act,
                                                 ptr(res), N, ptr(gamma), ptr(beta), None, None, 0, ptr(out), N, None, None, None)
                return L.fs2_op_conv_gemm(stream(), 0, ptr(A), K, ROWS, ptr(W), ptr(bias), taps, (taps - 1) // 2, K, N, act,
                                          ptr(res), N, None, None, 0, ptr(out), N)
            for _ in range(3):
                assert call() == 0, L.fs2_last_error(None)
            for mode in ("warm", "cold"):
                ts = []
                for _ in range(10):
                    if mode == "cold":
                        flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); call(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                us = float(np.median(ts))
                tf = 2.0 * ROWS * K * N * taps / us / 1e6
                line += f" | eng{eng} {mode} {us:7.1f} us {tf:6.1f} TF"
        print(line, flush=True)
