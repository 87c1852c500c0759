#!/usr/bin/env python3
"""Warp-stall samples of an ncu report per CUDA source line (needs -lineinfo and --import-source on).

    python scripts/ncu_src_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr, lines = None, None, []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":  # a source line (its SASS rows have an empty line number)
        d = dict(zip(hdr, r))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        lines.append((int(d["# Samples"] or 0), cur, int(r[0]), int(d["Instructions Executed"] or 0), stalls, r[1].strip()))
tot = sum(x[0] for x in lines)
print(f"total samples {tot}")
for s, f, ln, ex, st, src in sorted(lines, key=lambda x: -x[0])[:top_n]:
    top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:2])
    print(f"{100.0 * s / tot:5.1f}% {s:6d} exec={ex:>9} {f}:{ln:<5} [{top}] {src[:90]}")
