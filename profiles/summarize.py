#!/usr/bin/env python3
"""Turns gpurun_out/{launches_*.csv, *.ncu-rep} into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    python profiles/summarize.py kernel   gpurun_out/probe_r01.ncu-rep > profiles/r01_probe_kernel.md
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path):
    with open(path) as fh:
        txt = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(io.StringIO("".join(txt))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# launch list: {path}\n")
    print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
    print(f"total device time {tot / 1e6:.3f} ms over {sum(n for n, _ in agg.values())} launches\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.2f}% |")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, unit = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for val in rows[2:]:
        name = val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name.split('(')[0]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for i, h in enumerate(hdr):
            if h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                print(f"| {h} | {val[i]} | {unit[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
