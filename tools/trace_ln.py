"""Phase timestamps of the fused-LayerNorm epilogue (trace build: FS2_TRACE_BUILD=1): CTA 0, warp 2, first tile."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
for rows, K in [(26788, 256), (26788, 1024), (128, 256)]:
    A = torch.randn(rows, K, device=DEV); W = torch.randn(1, 256, K, device=DEV) / 16; bias = torch.randn(256, device=DEV)
    res = torch.randn(rows, 256, device=DEV); gm = torch.ones(256, device=DEV); bt = torch.zeros(256, device=DEV)
    out = torch.empty(rows, 256, device=DEV)
    for use_res in (True, False):
        call = lambda: L.fs2_op_conv_gemm_ln(stream(), ptr(A), K, rows, ptr(W), ptr(bias), 1, 0, K, 0, ptr(res) if use_res else None, 256,
                                             ptr(gm), ptr(bt), None, None, 0, ptr(out), 256, None, None, None)
        for _ in range(3): call()
        torch.cuda.synchronize()
        L.fs2_debug_set_flag(1, 1)
        call(); torch.cuda.synchronize()
        buf = (ctypes.c_int64 * 64)()
        L.fs2_debug_read_trace(buf, 64)
        L.fs2_debug_set_flag(1, 0)
        t = np.array(list(buf), dtype=np.int64); rel = (t - t[0]) / 1e3
        print(f"LN rows={rows} K={K} residual={use_res}: acc_ready {rel[4]:.2f} exit {rel[7]:.2f} | pass1 start {rel[31]:.2f} chunks " +
              " ".join(f"{x:.2f}" for x in rel[32:40]) + " | pass2 chunks " + " ".join(f"{x:.2f}" for x in rel[40:48]) +
              f" | residual wait {t[48]} cycles | store of sub-tile 2 (us since its start): wait_read {rel[50]-rel[49]:.3f} sts {rel[51]-rel[50]:.3f} "
              f"fence {rel[52]-rel[51]:.3f} tma_issue {rel[53]-rel[52]:.3f}; compute before it {rel[49]-rel[41]:.3f}")
