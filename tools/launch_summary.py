"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the last frame's sequence."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        seq.append((r[kn].split("(")[0], v, r[gs]))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, v, _ in seq:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-45s %4d %10.1f us  avg %8.1f" % (k, a[0], a[1], a[1] / a[0]))
    if len(sys.argv) > 2:
        idx = [i for i, s in enumerate(seq) if s[0].startswith("bilateral")]
        for s in seq[idx[-1]:]:
            print("%-40s %8.1f %s" % s)
