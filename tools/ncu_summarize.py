#!/usr/bin/env python3
"""Turn an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv`) into the two files
bench.py and the reviewers read:
  profiles/<tag>_kernels.csv   one line per captured kernel: time, DRAM bytes, pipe / memory utilisation, stalls
  profiles/ncu_traffic.json    {kernel name: dram_bytes per launch} for bench.py's roofline.traffic
usage: tools/ncu_summarize.py raw.csv <tag> <frames_per_launch> "<how it was captured>"
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected"]


def main():
    raw, tag, frames, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw, errors="ignore")))
    h, units = rows[0], rows[1]
    idx = {c: h.index(c) for c in COLS if c in h}
    ni, gi, bi = h.index("Kernel Name"), h.index("Grid Size"), h.index("Block Size")
    out = os.path.join(ROOT, "profiles", f"{tag}_kernels.csv")
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block"] + [f"{c} [{units[idx[c]]}]" for c in idx])
        for r in rows[2:]:
            name = re.sub(r"^.*::", "", r[ni].split("(")[0].replace("void ", ""))
            w.writerow([name, r[gi], r[bi]] + [r[idx[c]] for c in idx])
            def val(c):
                v = float(r[idx[c]].replace(",", "") or 0)
                u = units[idx[c]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            if "dram__bytes_read.sum" in idx:
                traffic.setdefault(name, {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                          "time_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "") or 0)})
    json.dump({"source": how, "frames_per_launch": frames, "kernels": traffic},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print("wrote", out, "and profiles/ncu_traffic.json with", len(traffic), "kernels")


if __name__ == "__main__":
    main()
