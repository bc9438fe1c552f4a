#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run in the build container).

    python scripts/ncu_summary.py launches gpurun_out/launches.csv      > profiles/launches_rNN.txt
    python scripts/ncu_summary.py raw      gpurun_out/prof.ncu-rep      > profiles/ncu_rNN.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}")
    print(f"{'kernel':58s} {'n':>4s} {'total ms':>10s} {'avg ms':>9s} {'share':>6s}")
    for k, a in agg.items():
        print(f"{k:58s} {a[0]:4d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e6:9.3f} {a[1] / tot:6.3f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: {path}")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')][:100]}   grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:62s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
