"""Per-CTA lifetimes of the attention kernel on the micro-benchmark input of tools/trace_attention.py (64 utterances of 640-704
frames, standalone launch) -- to compare with tools/trace_attention_ctas.py (the same kernel inside a config-2 forward).
Trace build (FS2_TRACE_BUILD=1)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
import fs2_b200
syn = fs2_b200.synthetic
for name, lens in (("uniform 640", [704] + [640] * 63), ("config-2 like (80..700)", [int(4.0 * n + 20) for n in syn.random_lengths(64, seed=0)])):
    gap = 12
    starts, r = [], gap
    for n in lens:
        starts.append(r); r += n + gap
    rows = r
    qkv = torch.randn(rows, 768, device=DEV)
    out = torch.zeros(rows, 256, device=DEV)
    ds = torch.tensor(starts, dtype=torch.int32, device=DEV); dl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    for _ in range(3):
        L.fs2_op_attention(stream(), ptr(qkv), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 2048 * 6
    buf = (ctypes.c_int64 * n)()
    L.fs2_debug_read_trace(buf, n)
    t = np.array(list(buf), dtype=np.int64).reshape(2048, 6)
    t = t[(t[:, 3] > 0) & (t[:, 5] > 0)]
    life = (t[:, 3] - t[:, 0]) / 1e3
    fit = np.polyfit(t[:, 5], life, 1)
    print(f"{name}: {len(t)} real CTAs, span {(t[:, 3].max() - t[:, 0].min()) / 1e3:.1f} us, lifetime ~= {fit[1]:.2f} us + {fit[0]:.3f} us per tile; "
          f"mean lifetime / tiles {np.mean(life / t[:, 5]):.3f} us")

# effective SM clock of CTA 0 during the launch: cycles (clock64 stamps of the per-tile trace) against nanoseconds (globaltimer)
lens = [704] + [640] * 63
gap = 12
starts, r = [], gap
for n in lens:
    starts.append(r); r += n + gap
rows = r
qkv = torch.randn(rows, 768, device=DEV)
out = torch.zeros(rows, 256, device=DEV)
ds = torch.tensor(starts, dtype=torch.int32, device=DEV); dl = torch.tensor(lens, dtype=torch.int32, device=DEV)
L.fs2_debug_set_flag(0, 4)
L.fs2_op_attention(stream(), ptr(qkv), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
torch.cuda.synchronize()
L.fs2_debug_set_flag(0, 0)
tr = out[starts[0]].cpu().numpy()[:176].reshape(11, 16)
buf = (ctypes.c_int64 * (2048 * 6))()
L.fs2_debug_read_trace(buf, 2048 * 6)
t = np.array(list(buf), dtype=np.int64).reshape(2048, 6)
cyc_first_s, cyc_last = tr[0, 1], tr[10, 3]
print(f"CTA 0: first S tile at {cyc_first_s:.0f} cycles, tile 10 done at {cyc_last:.0f} cycles; globaltimer: entry -> first S {t[0, 2] - t[0, 0]} ns, "
      f"entry -> exit {t[0, 3] - t[0, 0]} ns  =>  ~{cyc_last / max(t[0, 3] - t[0, 0], 1) * 1e3:.0f} MHz effective over the CTA's life")
per = np.diff(tr[:, 1])
print("CTA 0 cycles between S tiles:", " ".join(f"{int(x)}" for x in per))
