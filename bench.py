#!/usr/bin/env python
"""bench.py -- mel frames/s of the FastSpeech2 inference forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--math tf32|bf16|parity]

A step is one forward pass over one synthetic batch of BASELINE config 2 (64 utterances of 20-120 phonemes, mixed
speakers / emotions / arousal-valence, controls 1.0) per GPU; with N > 1 (launched by torchrun, one rank per GPU) every
rank runs its own copy of that batch (utterances are independent: no collective on the data path, weak scaling) and rank 0
prints ONE JSON line.  `value` is timed with inputs resident in HBM; `e2e` goes through the host-buffer entry
(`FastSpeech2B200.synthesize_host`: pinned H2D of the int64 inputs, forward, D2H of everything `synth_samples` reads --
the packed postnet mel rows, pitch, energy, durations, mel_lens) with the copies inside the timed region.

Beside the headline the line carries: `roofline` (dominant kernel against the measured burst peak), `kernel_ms_per_step`,
`hbm_kernels` (every memory-bound kernel at batch 512, where the data no longer fits L2), `latency` (config 1),
`bf16_mode`, `vocoder`, `config3_strong` (BASELINE config 3: ONE batch of 512 sharded over the N GPUs, by phonemes and
re-balanced by frames), `cpu_baseline`.  `--impl reference` times the CPU oracle port of the reference forward
(oracle/fs2_oracle.py; the reference itself is Python and /root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "mel frames/s (batch 64)"
UNIT = "frames/s"
CPU_SAMPLE_UTTS = 64      # the whole config-2 batch per CPU step (about 3-4 s on 16 cores)
CPU_SAMPLE_STEPS = 3      # cpu_baseline leg of the default run: about 10 s of CPU work
FLOPS_CONV9_PER_ROW = 2 * 9 * 256 * 1024
# dram bytes (read + written) of one dec.ffn_fused launch at batch 64, from the ncu --set full capture; None until captured
DRAM_TRAFFIC_FUSED = {"tf32": 46.13e6}   # 39.26 MB read + 6.87 MB written (profiles/r02d_ncu_full_forward_summary.csv, launch 36;
                                          # 44.11 MB in profiles/r02_ncu_full_ffn_fused_final_raw.csv: the written part varies with what L2 keeps)
NAMES = ("speakers", "emotions", "arousals", "valences", "texts", "src_lens")


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"],
                "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def measure_tf32_peak(dev):
    """Calibration only (cuBLAS, never on the product path): the burst TF32 rate of this GPU, measured the way
    MEASURED_PEAKS.json measures bf16 -- torch.matmul 8192^3, best of 10, CUDA events."""
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every 5 ms; the timed
    region of the default run lasts ~60 ms, shorter than one `nvidia-smi -lms` period), nvidia-smi as the fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms period"}
        try:   # one nvidia-smi query (not in the timed region: stated as such)
            q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                str(self.index)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(q[0]), "sm_max_mhz": float(q[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def oracle_cpu_run(sd, batch, n_utts, steps, warmup):
    """The CPU oracle port of the reference forward on the first n_utts utterances, fp32, all host threads."""
    from oracle import fs2_oracle as O
    sub = {k: (v[:n_utts] if torch.is_tensor(v) else v) for k, v in batch.items()}
    L = int(sub["src_lens"].max())
    sub["texts"] = sub["texts"][:, :L].contiguous()
    sub["max_src_len"] = L
    args = [sub[k] for k in NAMES]
    small = [a[:2] for a in args]
    L2 = int(small[5].max())
    small[4] = small[4][:, :L2].contiguous()
    for _ in range(max(warmup, 1)):
        O.forward(sd, *small, L2, loop_lr=True)
    times, frames = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        out = O.forward(sd, *args, L, loop_lr=True)
        times.append(time.perf_counter() - t0)
        frames = int(out[9].sum())
    return frames, times


def host_threads():
    """All the host cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which would silently turn the CPU
    arm into a single-thread run: set the torch thread count explicitly."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    host_threads()
    import fs2_b200
    syn = fs2_b200.synthetic
    sd = syn.synthetic_state_dict(seed=0)
    batch = syn.config2_batch(seed=0)
    cores = torch.get_num_threads()
    # bounded sample: calibrate on 8 utterances, then take the largest prefix of the batch that keeps the whole
    # --steps run under about 150 s on this host
    _, cal = oracle_cpu_run(sd, batch, 8, 1, 1)
    n_utts = CPU_SAMPLE_UTTS
    while n_utts > 8 and cal[0] * (n_utts / 8.0) * 1.3 * max(args.steps, 1) > 150.0:
        n_utts //= 2
    frames, times = oracle_cpu_run(sd, batch, n_utts, args.steps, args.warmup)
    total = sum(times)
    value = frames * len(times) / total
    sample = (f"first {n_utts} of the 64 config-2 utterances per step ({frames} frames), oracle port of "
              f"FastSpeech2.forward, fp32, torch CPU, Python-loop LengthRegulator")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: batch 64 per GPU, 20-120 phonemes, mixed speakers/emotions/arousal-valence, "
                                   "controls 1.0, random-init weights (seed 0)",
                       "timed_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def profile_forward(lib, model, dev_args, L, flush, runs, lead=1):
    """Per-kernel-class CUDA-event timing through the library's own profiling hooks (events bracket each launch on the
    launching stream, so the programmatic-dependent-launch overlap between kernels is serialised away: the classes sum
    to MORE than the step time).  Returns {label: (launches per forward, median ms per forward)}."""
    lib.fs2_profile_enable(model._ctx, 1)
    acc = {}
    for _ in range(runs):
        # `lead` L2 flushes queue GPU work ahead of the forward, so that the first small kernels of a stage are timed while
        # the host is already ahead of the device (an idle device would add the host's launch latency to their events)
        for _ in range(lead):
            flush.zero_()
        model(*dev_args, L)
        buf = (ctypes.c_char * 16384)()
        lib.fs2_profile_read(model._ctx, buf, 16384)
        for line in buf.value.decode().splitlines():
            label, n, ms = line.split()
            acc.setdefault(label, []).append((int(n), float(ms)))
    lib.fs2_profile_enable(model._ctx, 0)
    return {k: (v[0][0], float(np.median([x[1] for x in v]))) for k, v in acc.items()}


def hbm_kernel_table(prof, frames, phonemes, batch, t_max, peak_gbs):
    """Achieved GB/s of the memory-bound kernels from their ALGORITHMIC bytes (fp32 rows of 1 KB)."""
    P, F = phonemes, frames
    algo = {
        "embed_pe": 8 * P + 1024 * P,                         # ids in, rows out (embedding / PE tables are L2 resident)
        "add_cond": 2048 * P,                                 # read x, write x + spk + emo
        "bucket_embed_add": 2048 * P + 8 * P,                 # read x (+ raw prediction), write x + embedding (+ prediction)
        "length_regulator": 1024 * P + 8 * P + 1024 * F,      # one 1 KB row + cum + energy per phoneme in, 1 KB per frame out
        "unpack": 2 * 320 * F + (2 * 320 + 1) * batch * t_max,
    }
    out = {}
    for k, nbytes in algo.items():
        if k in prof and prof[k][1] > 0:
            ms = prof[k][1] / max(prof[k][0], 1)
            gbs = nbytes / (ms * 1e-3) / 1e9
            out[k] = {"bytes": nbytes, "us": ms * 1e3, "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / peak_gbs}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default=os.environ.get("FS2_MATH", "tf32"), choices=["tf32", "bf16", "parity"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline, e2e, roofline and latency only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    import fs2_b200
    from fs2_b200 import _lib, partition
    syn = fs2_b200.synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sd = syn.synthetic_state_dict(seed=0)
    jsons = syn.write_fixture_jsons(tempfile.mkdtemp(prefix="fs2_json_"))

    def new_model(math):
        m = fs2_b200.FastSpeech2B200(fs2_b200.config.default_preprocess_config(jsons), fs2_b200.config.default_model_config(),
                                     math_mode=math)
        m.load_state_dict(sd)
        return m.to(dev)

    model = new_model(args.math)
    tf32_peak = measure_tf32_peak(dev) if rank == 0 else None

    # weak scaling: every rank runs the SAME config-2 batch (seed 0), so the per-GPU work is identical by construction
    # and the max-over-ranks time measures the hardware, not the luck of a rank's length draw
    batch = syn.config2_batch(seed=0, batch=args.batch)
    dev_args = [batch[k].to(dev) for k in NAMES]
    host_batch = {k: batch[k].numpy() for k in NAMES}
    host_batch["max_src_len"] = batch["max_src_len"]
    L = batch["max_src_len"]

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, steps, sync_each=False):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for beg, end in evs:
            flush.zero_()                      # evict L2 between timed iterations (outside the event pair)
            if sync_each:
                barrier()                      # collectives inside the step: every rank starts the step together
            beg.record()
            step_fn()
            end.record()
        barrier()
        return [b.elapsed_time(e) for b, e in evs]

    out = None
    for _ in range(args.warmup):
        out = model(*dev_args, L)
        model.synthesize_host(host_batch, copy=False)
    torch.cuda.synchronize()
    frames = int(out[9].sum())
    mel_lens = out[9].tolist()
    launches = model.last_launch_count

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms = timed_loop(lambda: model(*dev_args, L), args.steps)
    clocks = sampler.stop()
    e2e_bytes = {}

    def e2e_loop(m, steps):
        """K host-buffer syntheses as a serving loop would issue them: batch i+1 is submitted (pinned H2D, forward) while
        the device->host reads of batch i are still landing, and batch i's results are taken right after.  Every step's
        H2D and D2H copies are inside the timed span: first submit -> last result on the host (device events on the
        launching stream and on the copy stream).  The L2 flush between steps is inside the span too (~45 us each)."""
        w0 = m.synthesize_host_async(host_batch, copy=False)     # untimed: both staging slots and the copy stream exist
        m.synthesize_host_async(host_batch, copy=False).wait()
        w0.wait()
        barrier()
        beg = torch.cuda.Event(enable_timing=True)
        beg.record()
        prev = None
        for i in range(steps):
            flush.zero_()
            h = m.synthesize_host_async(host_batch, copy=False)
            if prev is not None:
                prev.wait()
            prev = h
        _, _, h2d, d2h = prev.wait()
        e2e_bytes["h2d"], e2e_bytes["d2h"] = h2d, d2h
        span = beg.elapsed_time(prev.done)
        barrier()
        return span

    e2e_total = e2e_loop(model, args.steps)

    # the second arithmetic mode of the north star (bf16 operands), same workload, reported beside the headline
    other = None
    if args.math == "tf32" and not args.no_extras:
        m16 = new_model("bf16")
        for _ in range(args.warmup):
            o16 = m16(*dev_args, L)
            m16.synthesize_host(host_batch, copy=False)
        torch.cuda.synchronize()
        f16 = int(o16[9].sum())
        ms16 = timed_loop(lambda: m16(*dev_args, L), args.steps)
        other = (f16, float(sum(ms16)), float(e2e_loop(m16, args.steps)))
        del m16

    # the step after the path: HiFi-GAN generator on the mel this batch produced, reported beside
    # the headline (device-resident mel, L2 flushed); never part of `value`
    voc_line = None
    if world == 1 and not args.no_extras:
        try:
            mel_t, lens_t = out[1].transpose(1, 2), out[9]
            voc_line = {"workload": "HiFi-GAN V1 generator (hifigan/config.json) on the postnet mel of the same batch, "
                                    "random-init weights, ragged (frames beyond mel_lens skipped)"}
            for mode in ("tf32", "bf16"):
                voc = fs2_b200.HiFiGANGeneratorB200(math_mode=mode)
                voc.load_state_dict(syn.synthetic_vocoder_state_dict(0))
                voc = voc.to(dev)
                for _ in range(2):
                    voc(mel_t, mel_lens=lens_t)
                v = float(np.median(timed_loop(lambda: voc(mel_t, mel_lens=lens_t), min(args.steps, 5))))
                voc_line[mode] = {"ms_per_step": v, "mel_frames_per_s": frames / v * 1e3,
                                  "audio_samples_per_s": frames * 256 / v * 1e3,
                                  "x_realtime_22050hz": frames * 256 / 22050 / (v * 1e-3), "gpu_launches": voc.last_launch_count}
                del voc
                torch.cuda.empty_cache()
        except Exception as e:   # the vocoder is an extra: never let it take the headline line down
            voc_line = {"error": str(e)[:200]}

    # per-kernel-class CUDA-event timing (same workload, same process, after the timed region)
    lib = _lib.load_library()
    PROF_RUNS = 7
    prof = profile_forward(lib, model, dev_args, L, flush, PROF_RUNS)

    # the memory-bound kernels where they actually reach HBM: one batch of 512 (the batch-64 tensors live in the 126 MB L2)
    hbm = None
    c3 = None
    if not args.no_extras:
        big = syn.config2_batch(seed=0, batch=512)
        if world == 1:
            big_args = [big[k].to(dev) for k in NAMES]
            for _ in range(2):
                ob = model(*big_args, big["max_src_len"])
            torch.cuda.synchronize()
            prof_big = profile_forward(lib, model, big_args, big["max_src_len"], flush, 5, lead=12)
            hbm = {"workload": "config-3 batch (512 utterances) on one GPU, L2 flushed before each forward",
                   "kernels": hbm_kernel_table(prof_big, int(ob[9].sum()), int(big["src_lens"].sum()), 512, int(ob[0].shape[1]),
                                               measured_peaks()["hbm_gbs"])}
            del ob
        # ---- BASELINE config 3: ONE batch of 512 utterances sharded over the N GPUs (strong scaling)
        parts = partition.lpt_partition(big["src_lens"].tolist(), world)
        mine = partition.take(big, parts[rank])
        mine_args = [mine[k].to(dev) for k in NAMES]
        steps3 = max(3, min(args.steps, 10))
        for _ in range(2):
            o3 = model(*mine_args, mine["max_src_len"])
        torch.cuda.synchronize()
        t_plain = float(np.median(timed_loop(lambda: model(*mine_args, mine["max_src_len"]), steps3, sync_each=world > 1)))
        f_plain = float(o3[9].sum())
        t_reb, f_reb = None, None
        if world > 1:
            for _ in range(2):
                r3 = partition.rebalanced_forward(model, big, parts, rank)
            torch.cuda.synchronize()
            t_reb = float(np.median(timed_loop(lambda: partition.rebalanced_forward(model, big, parts, rank), steps3, sync_each=True)))
            f_reb = float(r3["mel_lens"].sum()) if r3["ids"] else 0.0
        # third arm: balance on frames PREDICTED before stage 1 by a speaking-rate prior over the conditioning, learned on
        # four OTHER batches (config-2 shape, seeds 11-14: every rank runs them untimed, so every rank holds the same
        # prior and plans the same shards without communicating)
        t_pri, f_pri = None, None
        if world > 1:
            prior = partition.RatePrior(len(syn.SPEAKERS), 5, 4, 5)
            for seed in (11, 12, 13, 14):
                ob = syn.config2_batch(seed=seed)
                oo = model(*[ob[k].to(dev) for k in NAMES], ob["max_src_len"])
                prior.observe({k: ob[k].numpy() for k in NAMES if k != "texts"}, oo[9].cpu().numpy())
            parts_p = partition.lpt_partition_by_prior({k: big[k].numpy() for k in NAMES if k != "texts"}, world, prior)
            mine_p = partition.take(big, parts_p[rank])
            mine_p_args = [mine_p[k].to(dev) for k in NAMES]
            for _ in range(2):
                p3 = model(*mine_p_args, mine_p["max_src_len"])
            torch.cuda.synchronize()
            t_pri = float(np.median(timed_loop(lambda: model(*mine_p_args, mine_p["max_src_len"]), steps3, sync_each=True)))
            f_pri = float(p3[9].sum())
        v = torch.tensor([t_plain, f_plain, t_reb or 0.0, f_reb or 0.0, t_pri or 0.0, f_pri or 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            vmax, vmin, vsum = v.clone(), v.clone(), v.clone()
            dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(vsum, op=dist.ReduceOp.SUM)
        else:
            vmax = vmin = vsum = v
        c3 = {"workload": "config3: ONE batch of 512 utterances (20-120 phonemes, seed 0) sharded over the GPUs; step = the "
                          "slowest rank (CUDA events, every rank starts each step at a barrier, L2 flushed)",
              "n_gpus": world, "total_frames": float(vsum[1]),
              "lpt_by_phonemes": {"ms_per_step": float(vmax[0]), "frames_per_s": float(vsum[1]) / float(vmax[0]) * 1e3,
                                  "frames_per_rank_min_max": [float(vmin[1]), float(vmax[1])],
                                  "note": "shards balanced on phoneme counts (all that is known before stage 1); no collective"}}
        if world > 1:
            c3["rebalanced_by_frames"] = {
                "ms_per_step": float(vmax[2]), "frames_per_s": float(vsum[3]) / float(vmax[2]) * 1e3,
                "frames_per_rank_min_max": [float(vmin[3]), float(vmax[3])],
                "note": "stage 1 on the phoneme shards, all-gather of mel_lens, LPT on the true stage-2 cost, one NCCL "
                        "all-to-all of the phoneme rows that change owner, stage 2 where the utterance landed "
                        "(partition.rebalanced_forward); the collectives and the host planning are inside the timed step"}

            c3["lpt_by_rate_prior"] = {
                "ms_per_step": float(vmax[4]), "frames_per_s": float(vsum[5]) / float(vmax[4]) * 1e3,
                "frames_per_rank_min_max": [float(vmin[5]), float(vmax[5])],
                "note": "shards balanced on frames predicted before stage 1 by partition.RatePrior (additive speaking-rate "
                        "model over speaker / emotion / arousal / valence, fitted on 256 utterances of four other batches); "
                        "no collective, no exchange"}

    # p50 single-utterance latency (BASELINE config 1), device-resident inputs, host sync included
    c1 = syn.config1_batch()
    c1_args = [c1[k].to(dev) for k in NAMES]
    lat = []
    for i in range(230):
        t0 = time.perf_counter()
        o1 = model(*c1_args, c1["max_src_len"])
        torch.cuda.synchronize()
        if i >= 30:
            lat.append((time.perf_counter() - t0) * 1e3)
    c1_frames = int(o1[9].sum())
    c1_launches = model.last_launch_count

    total_ms = float(sum(step_ms))
    o = other or (0, 0.0, 0.0)
    t = torch.tensor([total_ms, float(e2e_total), float(frames), o[1], o[2], float(o[0])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_total_ms, frames_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
        o = (float(tsum[5]), float(tmax[3]), float(tmax[4]))
    else:
        e2e_total_ms, frames_all = float(t[1]), float(frames)

    if rank == 0:
        peaks = measured_peaks()
        half = 0.5 if args.math in ("tf32", "parity") else 1.0
        # the dominant kernel is timed ALONE between two events (155 us): the burst peak is its denominator.  TF32 burst
        # peak: measured in this run with cuBLAS (calibration only); MEASURED_PEAKS.json's bf16 burst / 2 as the fallback
        burst = tf32_peak if (half == 0.5 and tf32_peak) else peaks["bf16_burst"] * half
        peak_source = ("torch.matmul TF32 8192^3 burst, measured in this run (calibration only)" if (half == 0.5 and tf32_peak)
                       else f"{peaks['source']} bf16_tflops (burst)" + (" / 2" if half == 0.5 else ""))
        # the decoder FFN: one fused launch per layer (conv9 -> ReLU -> w2 -> +x -> LayerNorm, stream-K over hidden
        # chunks) once the row tiles fill the machine, else conv9 and w2/LN as two launches with conv9 the dominant one
        fused = "dec.ffn_fused" in prof
        dom = "dec.ffn_fused" if fused else "dec.gemm_conv9"
        dom_what = ("decoder FFN in one launch: Conv1d k=9 256->1024 + ReLU + Conv1d k=1 1024->256 + residual + LayerNorm"
                    if fused else "decoder FFN Conv1d k=9 implicit GEMM, 256->1024")
        n_dom, ms_dom = prof.get(dom, (0, 0.0))
        per_launch_ms = ms_dom / max(n_dom, 1)
        terms = 3 if args.math == "parity" else 1
        flops_per_launch = (FLOPS_CONV9_PER_ROW + (2 * 1024 * 256 if fused else 0)) * frames
        # algorithmic bytes: activations in, weights once, and the output (hidden [rows,1024] for conv9 alone; the
        # fused kernel writes only the [rows,256] block output and re-reads x once as the residual)
        alg_bytes = (4 * (frames * 256 * 3 + 9 * 1024 * 256 + 1024 * 256) if fused
                     else 4 * (frames * 256 + 9 * 1024 * 256 + frames * 1024))
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the committed `ncu --set full`
        # captures (profiles/, batch 64): see profiles/README.md for which file each figure comes from
        traffic = None
        if args.batch == 64:
            traffic = DRAM_TRAFFIC_FUSED.get(args.math) if fused else {"tf32": 92.97e6, "bf16": 25.26e6}.get(args.math)
        achieved = flops_per_launch / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
        kernel_ms = {k: round(v[1], 4) for k, v in sorted(prof.items())}
        line = {
            "metric": METRIC, "value": frames_all * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.math, "data": "synthetic",
            "config": {"workload": f"config2: batch {args.batch} per GPU, 20-120 phonemes, mixed speakers/emotions/"
                                   "arousal-valence, controls 1.0, random-init weights (seed 0)",
                       "frames_per_step_per_gpu": frames, "phonemes_per_step_per_gpu": int(batch["src_lens"].sum()),
                       "l2": "256 MB buffer written between timed iterations (L2 flushed)",
                       "algorithmic_tflop_per_step": syn.algorithmic_flops(batch["src_lens"].tolist(), mel_lens) / 1e12},
            "e2e": {"value": frames_all * args.steps / (e2e_total_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": e2e_bytes.get("h2d", 0), "d2h_bytes_per_step": e2e_bytes.get("d2h", 0),
                    "ms_per_step": e2e_total_ms / args.steps,
                    "reads": "packed postnet mel rows, pitch, energy, log-duration, durations, mel_lens (utils/tools.py:228-243)",
                    "how": "synthesize_host_async: batch i+1 is submitted while batch i's device->host reads land (double-"
                           "buffered pinned staging, copy stream); span = first submit -> last result, copies and the "
                           "per-step L2 flush inside"},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": f"{dom} ({dom_what})",
                         "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": achieved / burst if burst else None,
                         "peak_source": peak_source,
                         "frac_of_half_measured_bf16_burst": achieved / (peaks["bf16_burst"] * half),
                         "tensor_flops_issued_per_launch": flops_per_launch * terms,
                         "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "per_launch_ms": per_launch_ms, "launches_per_step": n_dom,
                         "flops_per_launch": flops_per_launch},
            "kernel_ms_per_step": kernel_ms,
            "kernel_ms_note": "CUDA events around every launch: the programmatic-dependent-launch overlap between consecutive "
                              "kernels is serialised away, so the classes sum to more than ms_per_step",
            "latency": {"workload": "config1: single utterance, 16 phonemes, controls 1.0", "p50_ms": float(np.percentile(lat, 50)),
                        "p90_ms": float(np.percentile(lat, 90)), "frames": c1_frames, "gpu_launches": c1_launches,
                        "calls": len(lat)},
        }
        if hbm is not None:
            line["hbm_kernels"] = hbm
        if c3 is not None:
            line["config3_strong"] = c3
        if other is not None:
            line["bf16_mode"] = {"value": o[0] * args.steps / (o[1] * 1e-3), "e2e": o[0] * args.steps / (o[2] * 1e-3),
                                 "unit": UNIT, "ms_per_step": o[1] / args.steps,
                                 "note": "same workload with math_mode='bf16' (bf16 operands, fp32 accumulate/residual/outputs); "
                                         "tolerance stated in tests/test_gpu_bf16.py; not the headline"}
        if voc_line is not None:
            line["vocoder"] = voc_line
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            f_cpu, times = oracle_cpu_run(sd, syn.config2_batch(seed=0), CPU_SAMPLE_UTTS, CPU_SAMPLE_STEPS, 1)
            line["cpu_baseline"] = {"value": f_cpu * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{len(times)} forwards over the whole config-2 batch ({CPU_SAMPLE_UTTS} utterances, "
                                              f"{f_cpu} frames each, {sum(times):.1f} s in total), oracle port of "
                                              "FastSpeech2.forward, fp32, torch CPU, Python-loop LengthRegulator"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
