"""BASELINE config 3: ONE batch of 512 length-balanced utterances sharded over N GPUs (strong scaling), run under
torchrun on the GPU box:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config3.py [--gather]
Every rank takes its shard of `partition.lpt_partition` (identical on all ranks, no communication), runs the forward
on it and the job time is the max over ranks (CUDA events, L2 flushed).  --gather additionally all-gathers the postnet
mels over NCCL after the forward (reported separately: it is not part of the path)."""
import argparse, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import fs2_b200
from fs2_b200 import partition

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--gather", action="store_true")
ap.add_argument("--math", default="tf32")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
syn = fs2_b200.synthetic
m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(syn.write_fixture_jsons(tempfile.mkdtemp())),
                             fs2_b200.config.default_model_config(), math_mode=args.math)
m.load_state_dict(syn.synthetic_state_dict(0))
m = m.to(dev)
full = syn.config2_batch(seed=0, batch=args.batch)
parts = partition.lpt_partition(full["src_lens"].tolist(), world)
mine = partition.take(full, parts[rank])
names = ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")
dev_args = [mine[k].to(dev) for k in names]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(3):
    out = m(*dev_args, mine["max_src_len"])
torch.cuda.synchronize()
def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
ts, gs = [], []
barrier()
for _ in range(args.steps):
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    out = m(*dev_args, mine["max_src_len"])
    e1.record()
    if args.gather and world > 1:
        partition.gather_padded(out[1], out[9], parts[rank], args.batch)
    e2.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1)); gs.append(e1.elapsed_time(e2))
barrier()
t = torch.tensor([float(np.median(ts)), float(np.median(gs)), float(out[9].sum()), float(len(parts[rank]))], dtype=torch.float64, device=dev)
if world > 1:
    tmax, tsum, tmin = t.clone(), t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum, op=dist.ReduceOp.SUM); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
else:
    tmax = tsum = tmin = t
if rank == 0:
    print(json.dumps({"workload": f"config3: one batch of {args.batch} utterances (20-120 phonemes), LPT-sharded over {world} GPU(s), {args.math}",
                      "n_gpus": world, "ms_per_step_max_over_ranks": float(tmax[0]), "ms_fastest_rank": float(tmin[0]),
                      "frames_total": float(tsum[2]), "frames_per_s": float(tsum[2]) / float(tmax[0]) * 1e3,
                      "frames_per_rank_min_max": [float(tmin[2]), float(tmax[2])], "gather_ms": float(tmax[1]) if args.gather else None}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
