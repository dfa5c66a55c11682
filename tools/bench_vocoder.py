"""Time the HiFi-GAN generator on the B200 engine on the config-2 batch shape (run on the GPU box).
    python tools/bench_vocoder.py [--batch 64] [--steps 10]
Prints mel frames/s and audio samples/s (device-resident mel, L2 flushed between iterations)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fs2_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--frames", type=int, default=0, help="fixed frames per utterance (0 = the config-2 mel_lens)")
ap.add_argument("--math", default="tf32", choices=["tf32", "bf16"])
args = ap.parse_args()
dev = "cuda:0"
if os.environ.get("FS2_CL"):   # experiment: GEMM cluster size (2 = weight multicast across CTA pairs, 1 = independent CTAs)
    from fs2_b200 import _lib
    _lib.load_library().fs2_debug_set_flag(2, int(os.environ["FS2_CL"]))
syn = fs2_b200.synthetic
voc = fs2_b200.HiFiGANGeneratorB200(math_mode=args.math)
voc.load_state_dict(syn.synthetic_vocoder_state_dict(0))
voc = voc.to(dev)
if args.frames:
    lens = torch.full((args.batch,), args.frames, dtype=torch.int64)
else:
    # the mel lengths the acoustic model produces for the config-2 batch (synthetic weights, seed 0)
    import tempfile
    m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())),
                                 fs2_b200.config.default_model_config())
    m.load_state_dict(syn.synthetic_state_dict(0))
    m = m.to(dev)
    b = syn.config2_batch(seed=0, batch=args.batch)
    out = m(*[b[k].to(dev) for k in ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")], b["max_src_len"])
    lens = out[9].cpu()
    mel_bt = out[1]
T = int(lens.max())
if args.frames:
    mel_bt = torch.randn(args.batch, T, 80, device=dev) * 1.5 - 2.0
mel = mel_bt.transpose(1, 2)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(3):
    wav = voc(mel, mel_lens=lens)
torch.cuda.synchronize()
ts = []
for _ in range(args.steps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); wav = voc(mel, mel_lens=lens); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
frames = int(lens.sum())
print(json.dumps({"math": args.math, "workload": f"HiFi-GAN V1 generator, batch {args.batch}, {frames} mel frames ({frames * 256 / 22050:.1f} s of audio)",
                  "ms_per_step": ms, "mel_frames_per_s": frames / ms * 1e3, "samples_per_s": frames * 256 / ms * 1e3,
                  "x_realtime": frames * 256 / 22050 / (ms * 1e-3), "launches": voc.last_launch_count,
                  "algorithmic_tflop": frames * 0.6e9 / 1e12, "finite": bool(torch.isfinite(wav).all())}))
