import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, fs2_b200
from gpu_util import model_for, to_dev
syn = fs2_b200.synthetic
m = model_for(syn.synthetic_state_dict(0))
b = to_dev(syn.config1_batch())
args = [b[k] for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
for _ in range(20):
    m(*args, b["max_src_len"]); torch.cuda.synchronize()
os.environ["X"] = "1"
ts = []
for i in range(5):
    t0 = time.perf_counter(); o = m(*args, b["max_src_len"]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"python forward call {1e6*(t1-t0):.0f} us, tail sync {1e6*(t2-t1):.0f} us", file=sys.stderr)
