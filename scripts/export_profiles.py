#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches.csv into small committed summaries under profiles/ (named per round)."""
import csv
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def export_rep(path, name):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    extra = [i for i, h in enumerate(hdr) if ("tensor" in h or "tmem" in h.lower()) and i not in idx][:12]
    with open(os.path.join(OUT, f"{tag}_{name}_ncu_summary.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx + extra])
        w.writerow([units[i] for i in idx + extra])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx + extra])
    print("wrote", name)


def export_launches(path, name):
    rows = list(csv.reader(open(path)))
    rows = [r for r in rows if len(r) > 10 and r[0] != "ID"]
    agg = {}
    for r in rows:
        k = r[4]
        agg.setdefault(k, []).append(float(r[-1]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_{name}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(rows)} launches, cold-cache/serialised: compare SHARES\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'mean_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:60]:60s} {len(v):8d} {sum(v) / len(v) / 1e3:10.2f} {sum(v) / tot:7.3f}\n")
    print("wrote", name)


go = os.path.join(REPO, "gpurun_out")
for fn in sorted(os.listdir(go)):
    p = os.path.join(go, fn)
    stem = fn.rsplit(".", 1)[0]
    if stem.startswith(tag + "_"):      # files already named per round (r2_prof_*): no second prefix
        stem = stem[len(tag) + 1:]
    if fn.endswith(".ncu-rep"):
        export_rep(p, stem)
    elif "launches" in fn and fn.endswith(".csv"):
        export_launches(p, stem)
