#!/usr/bin/env python3
"""Rank the CUDA source lines of one kernel of an ncu report (captured with --import-source on, built with -lineinfo) by
executed warp instructions, with each line's share of the stall samples - the view DESIGN.md 4 "What the profiler said"
was read from (index divisions in the general epilogue, descriptor arithmetic in the MMA issue loops, the mask-bit loop of
the fused tail).
usage: tools/ncu_lines.py report.ncu-rep <kernel name substring> [top N]"""
import csv, io, subprocess, sys

rep, needle = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = path = None
agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] != "Line No" and fn and needle in fn and len(r) > 8 and r[2] == "-":      # the per-line aggregate row
        try:
            agg[(fn, path, int(r[0]))] = (int(r[7] or 0), int(r[6] or 0), r[1])
        except ValueError:
            pass
for kernel in sorted({k[0] for k in agg}):
    lines = {k[1:]: v for k, v in agg.items() if k[0] == kernel}
    tot = sum(v[0] for v in lines.values()) or 1
    samples = sum(v[1] for v in lines.values()) or 1
    print(f"{kernel[:110]}\n  executed warp instructions {tot}, stall samples {samples}")
    for (path, line), (inst, smp, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print(f"  {path}:{line:<5d} {100 * inst / tot:5.1f} % inst {100 * smp / samples:5.1f} % samples  {src.strip()[:120]}")
