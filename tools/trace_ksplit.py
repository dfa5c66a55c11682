"""Phase stamps of the K-split reduction (trace build): CTA 0 = rank 0 of the first cluster."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
for rows, K, N, taps in [(56, 256, 1024, 9), (56, 512, 512, 5)]:
    A = torch.randn(rows, K, device=DEV); W = torch.randn(taps, N, K, device=DEV) / 16; bias = torch.randn(N, device=DEV)
    out = torch.empty(rows, N, device=DEV)
    call = lambda: L.fs2_op_conv_gemm(stream(), 0, ptr(A), K, rows, ptr(W), ptr(bias), taps, (taps - 1) // 2, K, N, 0, None, N, None, None, 0, ptr(out), N)
    for _ in range(3): call()
    torch.cuda.synchronize()
    L.fs2_debug_set_flag(1, 1)
    call(); torch.cuda.synchronize()
    buf = (ctypes.c_int64 * 64)()
    L.fs2_debug_read_trace(buf, 64)
    L.fs2_debug_set_flag(1, 0)
    t = np.array(list(buf), dtype=np.int64); rel = (t - t[0]) / 1e3
    print(f"rows={rows} K={K} N={N} taps={taps}: acc_ready {rel[4]:.2f} | wait begin {rel[54]:.2f} partials arrived {rel[55]:.2f} first sub-tile summed {rel[56]:.2f} | epi_done {rel[5]:.2f} exit {rel[7]:.2f}")
