"""Per-tile timeline of CTA 0 for an A-resident vocoder conv (trace build): mma_begin, acc_free, epi_begin, epi_end (us)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
rows = 6_800_000
for K, N, taps, dil, use_res in [(32, 32, 11, 5, False), (32, 32, 11, 1, True), (32, 32, 3, 1, False), (64, 64, 11, 5, False)]:
    r = rows if K == 32 else rows // 2
    A = torch.randn(r, K, device=DEV); W = torch.randn(taps, N, K, device=DEV) / 16; bias = torch.zeros(N, device=DEV)
    res = torch.randn(r, N, device=DEV) if use_res else None
    out = torch.empty(r, N, device=DEV)
    call = lambda: L.fs2_op_conv_gemm_ex(stream(), ptr(A), K, r, ptr(W), ptr(bias), taps, dil, K, N, 3, 0.1, ptr(res), N, 1 if use_res else 0, 0, None, None, 0, 0, ptr(out), N)
    for _ in range(2): call()
    torch.cuda.synchronize()
    L.fs2_debug_set_flag(1, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_int64 * 40)()
    L.fs2_debug_read_trace(buf, 40)
    L.fs2_debug_set_flag(1, 0)
    t = np.array(list(buf), dtype=np.int64); rel = (t - t[0]) / 1e3
    print(f"K={K} N={N} taps={taps} dil={dil} res={use_res}: {e0.elapsed_time(e1)*1e3:.0f} us total, {r/128/148:.0f} tiles per SM")
    for j in range(6):
        print("   tile", j, " ".join(f"{x:8.2f}" for x in rel[8 + 4 * j: 12 + 4 * j]))
