"""Per-CTA and per-tile view of the persistent attention kernels (attention_tcq.cuh by default; FS2_ATTN_PERSISTENT=2 for the
one-group attention_tcp.cuh, whose per-item stamps fill the last table) in the LAST decoder attention launch of a config-2
forward under sustained load (trace build): key tiles per CTA (balance of the snake deal), lifetime, cycles per key tile, and for
CTA 5 the per-tile timeline.  Column "P->iter end": one-group kernel = the O update of the previous tile; two-group kernel = how
long after the P hand-over the accumulate group has finished the tile."""
import os, sys, ctypes, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fs2_b200
from fs2_b200 import _lib
syn = fs2_b200.synthetic
dev = "cuda:0"
m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())),
                             fs2_b200.config.default_model_config())
m.load_state_dict(syn.synthetic_state_dict(0))
m = m.to(dev)
b = syn.config2_batch(seed=0)
args = [b[k].to(dev) for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
for _ in range(int(os.environ.get("WARM", "200"))):
    out = m(*args, b["max_src_len"])
torch.cuda.synchronize()
lib = _lib.load_library()
n = 2048 * 6
buf = (ctypes.c_int64 * n)()
assert lib.fs2_debug_read_trace(buf, n) == 0
t = np.array(list(buf), dtype=np.int64).reshape(2048, 6)[:148]
ent, dep, first, ex, sm, tiles = (t[:, i] for i in range(6))
cyc, sm = sm >> 16, sm & 0xFFFF
t0 = ent.min()
life = (ex - ent) / 1e3
print(f"CTAs {len(t)}, kernel span {(ex.max() - t0) / 1e3:.1f} us; key tiles per CTA min/mean/max {tiles.min()} / {tiles.mean():.1f} / {tiles.max()} (sum {tiles.sum()})")
print(f"lifetime min/mean/max {life.min():.1f} / {life.mean():.1f} / {life.max():.1f} us; entry -> dependency wait done {((dep - ent) / 1e3).mean():.2f} us; wait done -> first S {((first - dep) / 1e3).mean():.2f} us")
print(f"effective clock {(cyc / np.maximum(ex - ent, 1) * 1e3).mean():.0f} MHz; cycles per key tile over the CTA's life: mean {(cyc / tiles).mean():.0f}, min {(cyc / tiles).min():.0f}, max {(cyc / tiles).max():.0f}")
work = (ex - dep) / 1e3
print(f"(exit - wait done) per key tile: mean {(work / tiles).mean():.3f} us; slowest CTA: {tiles[np.argmax(ex)]} tiles, exits at {(ex.max() - t0) / 1e3:.1f} us; earliest exit {(ex.min() - t0) / 1e3:.1f} us")

# per-tile timeline of CTA 5 (cycles relative to its first stamp)
buf2 = (ctypes.c_int64 * 513)()
assert lib.fs2_debug_read_trace(buf2, 513) == 0
tt = np.array(list(buf2)[:512], dtype=np.int64).reshape(64, 8)
n_t = int((tt[:, 1] > 0).sum())
base = tt[0, 0]
print("tile | softmax: wait S | S->P handed | P->iter end | period || QK issuer: wait | issue || PV operands ready (rel. to S ready)")
prev = None
for g in range(min(n_t, 40)):
    r = tt[g]
    per = r[1] - prev if prev is not None else 0
    prev = r[1]
    print(f"{g:3d} | {r[1]-r[0]:6d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {per:6d} || {r[5]-r[4]:6d} | {r[6]-r[5]:6d} || {r[7]-r[1]:6d}")
per = np.diff(tt[:n_t, 1])
print(f"S-to-S period over {n_t} tiles: mean {per.mean():.0f}, median {np.median(per):.0f} cycles")

buf3 = (ctypes.c_int64 * 130)()
assert lib.fs2_debug_read_trace(buf3, 130) == 0
it = np.array(list(buf3)[:128], dtype=np.int64).reshape(16, 8)
print("item | move_q of the next item: wait+copy | loop exit -> O there | first half FMAs | send(0) | second half | send(1) | tail total")
for k in range(16):
    r = it[k]
    if r[0] == 0: break
    print(f"{k:3d} | {r[7]-r[6] if r[6] else 0:6d} | {r[1]-r[0]:6d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {r[4]-r[3]:6d} | {r[5]-r[4]:6d} | {r[5]-r[0]:6d}")
