import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, fs2_b200
from fs2_b200 import _lib
from gpu_util import model_for, to_dev
syn = fs2_b200.synthetic
m = model_for(syn.synthetic_state_dict(0))
b = to_dev(syn.config1_batch())
args = [b[k] for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
for _ in range(20):
    m(*args, b["max_src_len"]); torch.cuda.synchronize()
lib = _lib.load_library()
lib.fs2_profile_enable(m._ctx, 1)
m(*args, b["max_src_len"]); torch.cuda.synchronize()
buf = (ctypes.c_char * 8192)()
lib.fs2_profile_read(m._ctx, buf, 8192)
lib.fs2_profile_enable(m._ctx, 0)
tot = 0
for line in buf.value.decode().splitlines():
    label, n, ms = line.split(); tot += float(ms)
    print(f"{label:24s} n={n:>3s} total {float(ms)*1e3:8.1f} us  per launch {float(ms)*1e3/int(n):6.1f} us")
print("sum of labelled kernel time: %.1f us" % (tot * 1e3))
ts = []
for _ in range(50):
    t0 = time.perf_counter(); m(*args, b["max_src_len"]); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("p50 latency %.1f us" % (sorted(ts)[25] * 1e6))
