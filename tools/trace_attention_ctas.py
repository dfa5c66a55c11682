"""Per-CTA timeline of the LAST decoder attention launch of a config-2 forward (trace build: FS2_TRACE_BUILD=1):
entry / after griddepcontrol.wait / first S tile / exit of every CTA and the SM it ran on -> how long a CTA lives per
key tile, how long an SM waits between two CTAs, how busy the SMs are."""
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
for _ in range(int(os.environ.get("WARM", "200"))):   # sustained load: the SM clock under the bench's conditions
    out = m(*args, b["max_src_len"])
torch.cuda.synchronize()
lib = _lib.load_library()
n = 2048 * 6
buf = (ctypes.c_int64 * n)()
assert lib.fs2_debug_read_trace(buf, n) == 0
t = np.array(list(buf), dtype=np.int64).reshape(2048, 6)
t = t[t[:, 3] > 0]
t0 = t[:, 0].min()
ent, dep, first, ex, sm, tiles = (t[:, i] for i in range(6))
cyc, sm = sm >> 16, sm & 0xFFFF
t[:, 4] = sm
print(f"CTAs {len(t)} (real: {(tiles > 0).sum()}), kernel span {(ex.max() - t0) / 1e3:.1f} us, SMs used {len(np.unique(sm))}")
real = tiles > 0
life = (ex - ent)[real] / 1e3
print(f"real CTA lifetime: mean {life.mean():.2f} us, per key tile (lifetime / tiles) mean {(life / tiles[real]).mean():.2f} us")
print(f"entry -> first S: mean {((first - ent)[real] / 1e3).mean():.2f} us; entry -> dependency wait done: {((dep - ent)[real] / 1e3).mean():.2f} us")
mhz = cyc[real] / np.maximum((ex - ent)[real], 1) * 1e3
print(f"effective SM clock over the CTAs' lives (clock64 cycles / globaltimer ns): mean {mhz.mean():.0f} MHz, min {mhz.min():.0f}, max {mhz.max():.0f}")
fitc = np.polyfit(tiles[real], cyc[real], 1)
print(f"lifetime ~= {fitc[1]:.0f} cycles + {fitc[0]:.0f} cycles per 64-key tile")
fit = np.polyfit(tiles[real], life, 1)
print(f"lifetime ~= {fit[1]:.2f} us + {fit[0]:.3f} us per 64-key tile (least squares over the real CTAs)")
gaps, busy = [], []
for s_ in np.unique(sm):
    rows = t[sm == s_]
    rows = rows[np.argsort(rows[:, 0])]
    gaps += ((rows[1:, 0] - rows[:-1, 3]) / 1e3).tolist()
    busy.append((rows[:, 3] - rows[:, 0]).sum() / (ex.max() - t0))
print(f"gap between exit of a CTA and entry of the next one on the same SM: mean {np.mean(gaps):.2f} us, p90 {np.percentile(gaps, 90):.2f} us")
print(f"fraction of the kernel span with a CTA resident per SM: mean {np.mean(busy):.2f}, min {np.min(busy):.2f}")
print(f"last CTA entry at {(ent.max() - t0) / 1e3:.1f} us; sum of tiles {tiles.sum()}")
