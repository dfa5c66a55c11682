import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from gpu_util import DEV, lib, ptr, stream
L = lib()
lens = [704] + [640] * 63      # utterance 0 is the longest: block 0 of the longest-first work list takes its first query tile
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
e0.record(); L.fs2_op_attention(stream(), ptr(qkv), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out)); e1.record(); torch.cuda.synchronize()
print("attention 64x640: %.1f us" % (e0.elapsed_time(e1) * 1e3), "CTAs", 64 * 2 * 5)
L.fs2_debug_set_flag(0, 4)
out.zero_()
L.fs2_op_attention(stream(), ptr(qkv), rows, ptr(ds), ptr(dl), len(lens), max(lens), ptr(out))
torch.cuda.synchronize()
L.fs2_debug_set_flag(0, 0)
t = out[starts[0]].cpu().numpy()[:176].reshape(11, 16)
print("2-SM kernel (FS2_ATTN_PAIR unset): tile of 128 keys; old kernel (FS2_ATTN_PAIR=0): 64 keys")
print("tile | sm: wait_s  s_ready  p_arrived acc_done | mma: before_p p_ready v_ready | qk(j): begin k_ready issued committed | pv(j): issued committed")
for j in range(11):
    print(j, " ".join(f"{int(x):8d}" for x in t[j, :14]))
