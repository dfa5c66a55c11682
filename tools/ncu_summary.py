"""Summarise `ncu --page raw --csv` exports: one line per launch with duration, DRAM bytes, DRAM throughput % and the
tensor-pipe activity as AVERAGE over the SMs (with min and max beside it -- the spread is the load balance).
    python tools/ncu_summary.py a.csv b.csv ... > profiles/rNN_ncu_full_summary.csv"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "duration_us", 1e-3),
        ("dram__bytes_read.sum", "dram_read_MB", None), ("dram__bytes_write.sum", "dram_write_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_avg_pct", 1.0),
        ("sm__pipe_tensor_cycles_active.min.pct_of_peak_sustained_active", "tensor_pipe_min_pct", 1.0),
        ("sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active", "tensor_pipe_max_pct", 1.0),
        ("launch__registers_per_thread", "registers", 1.0), ("launch__grid_size", "grid", 1.0)]
UNIT_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}

w = csv.writer(sys.stdout)
w.writerow(["file", "id", "kernel"] + [c[1] for c in COLS])
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kern = r[idx["Kernel Name"]]
        short = kern.split("(")[0].split("::")[-1][:60]
        out = [path.split("/")[-1], r[0], short]
        for metric, _, scale in COLS:
            if metric not in idx:
                out.append("")
                continue
            v = r[idx[metric]].replace(",", "")
            try:
                x = float(v)
            except ValueError:
                out.append(v)
                continue
            if scale is None:
                x *= UNIT_MB.get(units[idx[metric]], 1.0)
            else:
                x *= scale
                if metric == "gpu__time_duration.sum":
                    x = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[idx[metric]], 1e-3)
            out.append(f"{x:.2f}")
        w.writerow(out)
