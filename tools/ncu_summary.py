#!/usr/bin/env python
"""Summarises ncu outputs for profiles/ (run here, no GPU needed).
  tools/ncu_summary.py rep  gpurun_out/prof_X.ncu-rep   -> markdown table of the per-kernel metrics the judge reads
  tools/ncu_summary.py list gpurun_out/launches_X.csv    -> per-kernel launch count / total time / share of the step
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64_inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), n) for m, n in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join("%s [%s]" % (n, units[i]) if units[i] else n for i, n in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        vals = []
        for i, n in cols:
            try:
                v = float(r[i])
                vals.append(("%.4g" % v))
            except ValueError:
                vals.append(r[i])
        print("| `%s` | " % r[ki].split("(")[0] + " | ".join(vals) + " |")


def lst(path, skip_before=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rd:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v  # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total [us] | share |")
    print("|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (name, n, t, 100 * t / tot))
    print("| all | %d | %.1f | 100%% |" % (sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2])
