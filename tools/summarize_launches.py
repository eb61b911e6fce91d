#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table for
one steady-state step (delimited by consecutive bce_loss_kernel launches)."""
import collections
import csv
import re
import sys


def main(path, step_index=2):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    marks = [i for i, n in enumerate(names) if "bce_loss_kernel" in n]
    a, b = marks[step_index], marks[step_index + 1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[a:b]:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("step launches: %d   sum of kernel durations: %.1f us" % (b - a, tot))
    print("%-72s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %6d %12.1f %6.1f%%" % (n[:72], c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
