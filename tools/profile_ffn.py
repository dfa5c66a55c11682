"""The fused FFN kernel on a batch-512-sized row count, for `ncu -k regex:ffn_fused_kernel` (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 206000
g = torch.Generator().manual_seed(0)
x = torch.randn(rows, 256, generator=g).to(DEV)
w1 = (torch.randn(9, 1024, 256, generator=g) / np.sqrt(2304)).to(DEV); b1 = torch.zeros(1024, device=DEV)
w2 = (torch.randn(256, 1024, generator=g) / 32).to(DEV); b2 = torch.zeros(256, device=DEV)
gm, bt = torch.ones(256, device=DEV), torch.zeros(256, device=DEV)
out = torch.empty(rows, 256, device=DEV)
L = lib()
for i in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert L.fs2_op_ffn_fused(stream(), ptr(x), rows, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(gm), ptr(bt), None, None, 0, ptr(out)) == 0
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"ffn_fused rows={rows}: {ms:.3f} ms, {rows * 2 * (2304 * 1024 + 1024 * 256) / ms / 1e9:.1f} TFLOP/s")
