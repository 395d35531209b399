#!/usr/bin/env python
"""Turns ncu outputs brought back from the GPU box into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r01_stage1.csv > profiles/...
    python tools/summarize_ncu.py full gpurun_out/prof_xxx.ncu-rep > profiles/...
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("hi3d::", "")[:60]
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n"
          f"# {n} launches, {tot:.3f} ms total")
    print(f"{'kernel':62s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:62s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot * 100:6.1f}% {v[1] / v[0] * 1e3:9.1f}")


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_utchmma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path):
    if path.endswith(".csv"):        # already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    extra = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_elapsed" in h and "realtime" in h]
    idx += [(w, hdr.index(w)) for w in extra[:2]]
    print(f"# ncu --set full --clock-control none; source report {path.split('/')[-1]} (not committed: size)")
    for r in rows[2:]:
        print("---")
        for w, i in idx:
            print(f"{w:75s} {r[i][:70]:>30s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
