#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report (run here, no GPU needed).
  tools/ncu_source.py gpurun_out/prof_X.ncu-rep <kernel-regex> [top N]
Aggregates 'Instructions Executed' and stall samples per CUDA source line over all files (inlined headers included)."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, items = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            hdr = None
            continue
        if r[0] in ("Kernel Name", "Function Name"):
            continue
        if r[0] in ("Line No", "#"):
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].strip():
            continue  # SASS rows (empty line number) are already aggregated into their source line
        d = {}
        for kk, vv in zip(hdr, r):
            d.setdefault(kk, vv)
        try:
            inst = float(d.get("Instructions Executed", "0") or 0)
            samp = float(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        if inst or samp:
            items.append((inst, samp, fname, d[hdr[0]], d.get("Source", "")[:110]))
    ti, ts = sum(i[0] for i in items), sum(i[1] for i in items)
    print("total warp instructions %.4g, samples %d" % (ti, ts))
    print("by file:")
    files = {}
    for i in items:
        f = files.setdefault(i[2], [0, 0])
        f[0] += i[0]
        f[1] += i[1]
    for f, (a, b) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print("  %-20s inst %5.1f%%  samples %5.1f%%" % (f, 100 * a / max(ti, 1), 100 * b / max(ts, 1)))
    if "--by-inst" in sys.argv:
        print("top lines by instructions:")
        for inst, samp, f, ln, src in sorted(items, key=lambda t: -t[0])[:top]:
            print("  %5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * inst / max(ti, 1), 100 * samp / max(ts, 1), f, ln, src.strip()))
        return
    print("top lines by samples:")
    for inst, samp, f, ln, src in sorted(items, key=lambda t: -t[1])[:top]:
        print("  %5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * samp / max(ts, 1), 100 * inst / max(ti, 1), f, ln, src.strip()))


if __name__ == "__main__":
    main()
