"""How much would CUDA graphs save?  Capture stage 2 (LR + decoder + mel + postnet) with torch's
graph capture around the C-ABI call and compare eager vs replay."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, fs2_b200
from fs2_b200 import _lib
from gpu_util import model_for, to_dev
syn = fs2_b200.synthetic
m = model_for(syn.synthetic_state_dict(0))
lib = _lib.load_library()
for name, batch in (("config1", syn.config1_batch()), ("config2", syn.config2_batch(0))):
    b = to_dev(batch)
    args = [b[k] for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
    for _ in range(5):
        out = m(*args, b["max_src_len"])
    torch.cuda.synchronize()
    mel, post, mask = torch.empty_like(out[0]), torch.empty_like(out[1]), torch.empty_like(out[7])
    io = _lib.Stage2IO(mel=mel.data_ptr(), postnet=post.data_ptr(), mel_mask=mask.data_ptr())
    def eager():
        s = torch.cuda.current_stream().cuda_stream
        assert lib.fs2_forward_stage2(m._ctx, s, C.byref(io)) == 0
    def timeit(fn, n=50):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_eager = timeit(eager)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eager()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        eager()
    t_graph = timeit(g.replay)
    ok = torch.equal(post, out[1])
    print(f"{name}: stage 2 eager {t_eager:.1f} us, graph replay {t_graph:.1f} us, outputs equal {ok}")
