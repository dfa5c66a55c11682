"""One A-resident vocoder conv (32 channels, k = 11, dilation 5, 6.8 M rows) for `ncu -k regex:conv_gemm_tc2_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
r, K, N, taps, dil = 6_800_000, 32, 32, 11, 5
A = torch.randn(r, K, device=DEV); W = torch.randn(taps, N, K, device=DEV) / 16; bias = torch.zeros(N, device=DEV)
out = torch.empty(r, N, device=DEV)
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert L.fs2_op_conv_gemm_ex(stream(), ptr(A), K, r, ptr(W), ptr(bias), taps, dil, K, N, 3, 0.1, None, N, 0, 0, None, None, 0, 0, ptr(out), N) == 0
    e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1)*1e3:.0f} us; algorithmic bytes {2*r*K*4/1e6:.0f} MB")
