#!/usr/bin/env python3
"""Rank the SASS instructions of one kernel of an ncu report by stall samples.
usage: tools/ncu_hot.py report.ncu-rep <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) == len(h) and r[isamp] != "# Samples"]
print("total samples", sum(int(r[isamp] or 0) for r in data), "static instructions", len(data),
      "executed (warp-level)", sum(int(r[iex] or 0) for r in data))
agg = {}
for r in data:
    for i in stall:
        v = int(r[i] or 0)
        if v: agg[h[i]] = agg.get(h[i], 0) + v
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:top_n]:
    st = sorted(((h[i], int(r[i] or 0)) for i in stall if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:2]
    print(r[ia][-5:], r[isamp].rjust(6), r[iex].rjust(8), r[isrc][:64].ljust(64), st)
