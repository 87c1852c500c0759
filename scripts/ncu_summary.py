#!/usr/bin/env python3
"""Summarise ncu artefacts brought back from the GPU box into profiles/ (tracked).

  python scripts/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep [more.ncu-rep] > profiles/rNN_ncu_full.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# per-kernel device time from {path} (ncu --metrics gpu__time_duration.sum, cold-cache, serialised)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n:70s} n={c:4d} total={t / 1e6:9.3f} ms share={t / tot * 100:5.1f}% avg={t / c / 1e3:9.1f} us")


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(out.splitlines()))
        hdr, units, rows = r[0], r[1], r[2:]
        print(f"# {p}")
        for row in rows:
            print("kernel:", row[hdr.index("Kernel Name")])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"  {w:75s} {row[i]} {units[i]}")
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
