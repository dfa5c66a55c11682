"""One config-2 forward bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` (run on the GPU box).
    python tools/profile_target.py [tf32|bf16]
Launch order inside the bracket (64 launches at config 2, TF32; ncu IDs): 0 layout_scan 1 row_meta (+ source mask, attention work list, init)
2 embed_pe 3 cond | 4-23 encoder (4 x [qkv, attention, fc+LN, conv9, w2+LN]; the last w2+LN adds the conditioning vectors) |
24-25 duration + pitch predictors (one launch per layer for both) 26 bucket+embed 27-28 energy predictor 29 durations 30 layout_scan |
31 row_meta 32 length regulator (+ energy add, PE) | 33-56 decoder (6 x [qkv, persistent attention, fc+LN, fused FFN]) |
57 mel_linear 58-62 PostNet 63 unpack."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fs2_b200
math = sys.argv[1] if len(sys.argv) > 1 else "tf32"
syn = fs2_b200.synthetic
dev = "cuda:0"
m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())),
                             fs2_b200.config.default_model_config(), math_mode=math)
m.load_state_dict(syn.synthetic_state_dict(0))
m = m.to(dev)
b = syn.config2_batch(seed=0)
args = [b[k].to(dev) for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")]
for _ in range(3):
    out = m(*args, b["max_src_len"])
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m(*args, b["max_src_len"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("frames", int(out[9].sum()), "launches", m.last_launch_count, "math", math)
