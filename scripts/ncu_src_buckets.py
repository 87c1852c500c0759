#!/usr/bin/env python3
"""Bucket the warp-stall samples of an ncu source page by basic block (runs of equal execution count).

    python scripts/ncu_src_buckets.py gpurun_out/prof.ncu-rep [min_fraction] [launch_index]
"""
import csv, subprocess, sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
skip = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *skip], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
si, ai, ei, ad = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address")
data = rows[h + 1:]
for cut, r in enumerate(data):  # a report may hold several launches: keep the first one
    if r and r[0] == "Kernel Name":
        data = data[:cut]
        break
data = [r for r in data if len(r) > max(si, ai, ei, ad)]
tot = sum(int(r[si]) for r in data)
print(f"{rows[0][1][:90] if len(rows[0]) > 1 else ''}\ntotal samples {tot}, {len(data)} instructions")
# who branches to each spin loop (labels mbarrier waits by their call site)
addr_of = {r[ad]: i for i, r in enumerate(data)}
i = 0
while i < len(data):
    j, e, s = i, data[i][ei], 0
    stall = {}
    while j < len(data) and data[j][ei] == e:
        s += int(data[j][si])
        for c in range(len(hdr)):
            if hdr[c].startswith("stall_") and "Not Issued" not in hdr[c] and data[j][c].isdigit():
                stall[hdr[c]] = stall.get(hdr[c], 0) + int(data[j][c])
        j += 1
    if s > tot * thr:
        top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(stall.items(), key=lambda x: -x[1])[:2])
        src = ""
        if "TRYWAIT" in " ".join(r[ai] for r in data[i:j]):
            tgt = data[i][ad][2:]
            callers = [k for k, r in enumerate(data) if tgt in r[ai] and not (i <= k < j)]
            src = f"  <- wait loop entered from instr {callers}"
        print(f"{i:5d}-{j - 1:5d} exec={e:>10} samples={s:7d} ({100.0 * s / tot:4.1f}%) [{top}] {data[i][ai].strip()[:50]}{src}")
    i = j
