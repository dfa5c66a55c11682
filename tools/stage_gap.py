"""Where the host sits between the two stages of a config-2 forward: time from the return of the stage-1 call (host synced on the
frame counts) to the return of the stage-2 call (all launches enqueued), and the device-side idle gap between the last stage-1
kernel and the first stage-2 kernel (CUDA events recorded right before / after the calls)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fs2_b200
syn = fs2_b200.synthetic
dev = "cuda:0"
m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())),
                             fs2_b200.config.default_model_config())
m.load_state_dict(syn.synthetic_state_dict(0)); m = m.to(dev)
b = syn.config2_batch(seed=0)
args = [b[k].to(dev) for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
s1, s2 = m._stage1, m._stage2
rec = {}
def stage1(*a, **k):
    rec["e0"] = torch.cuda.Event(enable_timing=True); rec["e0"].record()
    r = s1(*a, **k)
    rec["t1"] = time.perf_counter()
    rec["e1"] = torch.cuda.Event(enable_timing=True); rec["e1"].record()     # device idle here: lands immediately
    return r
def stage2(*a, **k):
    rec["t2a"] = time.perf_counter()
    r = s2(*a, **k)
    rec["t2b"] = time.perf_counter()
    rec["e2"] = torch.cuda.Event(enable_timing=True); rec["e2"].record()
    return r
m._stage1, m._stage2 = stage1, stage2
rows = []
for i in range(30):
    t0 = time.perf_counter()
    out = m(*args, b["max_src_len"])
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if i >= 10:
        rows.append(((rec["t1"] - t0) * 1e3, (rec["t2a"] - rec["t1"]) * 1e3, (rec["t2b"] - rec["t2a"]) * 1e3, (t3 - rec["t2b"]) * 1e3,
                     rec["e0"].elapsed_time(rec["e1"]), rec["e1"].elapsed_time(rec["e2"]), (t3 - t0) * 1e3))
r = np.median(np.array(rows), axis=0)
print(f"host: stage-1 call {r[0]:.3f} ms | between the calls {r[1]:.3f} | stage-2 call (enqueue) {r[2]:.3f} | wait for the device {r[3]:.3f} | total {r[6]:.3f}")
print(f"device: stage 1 (event to event) {r[4]:.3f} ms | stage 2 incl. the idle gap before its first kernel {r[5]:.3f} ms")
