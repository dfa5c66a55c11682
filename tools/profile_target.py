"""One config-2 forward bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off` (run on the GPU box).
    python tools/profile_target.py [tf32|bf16]
Launch order inside the bracket (75 launches): 1 layout_scan 2 row_meta 3 src_mask 4 embed_pe | 5-24 encoder (4 x [qkv,
attention, fc+LN, conv9, w2+LN]) | 25 cond 26 add_cond | 27-28 duration predictor 29-30 pitch predictor 31 bucket+embed
32-33 energy predictor 34 bucket+embed 35 durations 36 layout_scan | 37 row_meta 38 length regulator | 39-68 decoder (6 x 5)
| 69 mel_linear 70-74 PostNet 75 unpack."""
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
