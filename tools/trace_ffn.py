"""Phase timestamps of the fused FFN kernel (trace build: FS2_TRACE_BUILD=1): cluster 0, its first units.
Per unit: issuer start -> conv issued -> hidden ready (ReLU handed back) -> GEMM2 issued; epilogue warp 0: conv complete ->
ReLU done -> out complete -> segment epilogue done.  Times in microseconds since the first stamp."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 26788
x = torch.randn(rows, 256, device=DEV); w1 = torch.randn(9, 1024, 256, device=DEV) / 48; b1 = torch.randn(1024, device=DEV)
w2 = torch.randn(1, 256, 1024, device=DEV) / 32; b2 = torch.randn(256, device=DEV)
gm = torch.ones(256, device=DEV); bt = torch.zeros(256, device=DEV); y = torch.empty(rows, 256, device=DEV)
call = lambda: L.fs2_op_ffn_fused(stream(), ptr(x), rows, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(gm), ptr(bt), None, None, 0, ptr(y))
for _ in range(3): call()
torch.cuda.synchronize()
L.fs2_debug_set_flag(1, 1)
call(); torch.cuda.synchronize()
buf = (ctypes.c_int64 * 64)()
L.fs2_debug_read_trace(buf, 64)
L.fs2_debug_set_flag(1, 0)
t = np.array(list(buf), dtype=np.int64).reshape(8, 8)
t0 = t[0, 0]
names = ["start", "conv_issued", "hid_ready", "gemm2_issued", "conv_done", "relu_done", "out_done", "seg_done"]
print("unit " + " ".join(f"{n:>12s}" for n in names))
for u in range(8):
    print(f"{u:4d} " + " ".join(f"{(v - t0) / 1e3:12.2f}" if v else f"{'-':>12s}" for v in t[u]))

ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); call(); ev[1].record(); torch.cuda.synchronize()
print(f"one launch, events: {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
big = (ctypes.c_int64 * (256 * 6))()
L.fs2_debug_read_trace(big, 256 * 6)
c = np.array(list(big), dtype=np.int64).reshape(256, 6)
c = c[c[:, 0] > 0]
t0 = c[:, 0].min()
print("cta  entry  after_wait  exit  first_unit units smid   (us since the first CTA entered)")
for i, r in enumerate(c):
    if i % 2 == 0:
        print(f"{i:3d} {(r[0]-t0)/1e3:8.2f} {(r[1]-t0)/1e3:8.2f} {(r[2]-t0)/1e3:8.2f} {r[3]:5d} {r[4]:3d} {r[5]:4d}")
