import os, sys, time, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np, torch, fs2_b200
syn = fs2_b200.synthetic
dev = torch.device("cuda:0")
m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())), fs2_b200.config.default_model_config())
m.load_state_dict(syn.synthetic_state_dict(0)); m = m.to(dev)
b = syn.config2_batch(seed=0)
names = ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")
hb = {k: b[k].numpy() for k in names}; hb["max_src_len"] = b["max_src_len"]
for _ in range(4): m.synthesize_host(hb, copy=False)
torch.cuda.synchronize()
def sync_loop(n):
    t0 = time.perf_counter()
    for _ in range(n): m.synthesize_host(hb, copy=False)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def async_loop(n):
    t0 = time.perf_counter(); prev = None
    for _ in range(n):
        h = m.synthesize_host_async(hb, copy=False)
        if prev is not None: prev.wait()
        prev = h
    prev.wait(); torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("sync  ms/step", sync_loop(20), sync_loop(20))
print("async ms/step", async_loop(20), async_loop(20))
# where does the time go in one async call
for _ in range(3):
    t0 = time.perf_counter(); h = m.synthesize_host_async(hb, copy=False); t1 = time.perf_counter(); h.wait(); t2 = time.perf_counter()
    print(f"submit {1e3*(t1-t0):.2f} ms, wait {1e3*(t2-t1):.2f} ms")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
flush32 = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for name, f in (("uint8.zero_", lambda: flush.zero_()), ("float32.zero_", lambda: flush32.zero_()), ("float32.fill_(1)", lambda: flush32.fill_(1.0))):
    for _ in range(2): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(name, "256 MB:", e0.elapsed_time(e1) / 10, "ms")
def async_loop_flush(n):
    torch.cuda.synchronize(); t0 = time.perf_counter(); prev = None
    for _ in range(n):
        flush.zero_()
        h = m.synthesize_host_async(hb, copy=False)
        if prev is not None: prev.wait()
        prev = h
    prev.wait(); torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("async with flush ms/step", async_loop_flush(20), async_loop_flush(20))
