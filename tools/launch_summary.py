#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    t = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        t[name][0] += 1
        t[name][1] += v
    tot = sum(v[1] for v in t.values())
    print(f"{'kernel':34s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>6s}")
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:34s} {v[0]:5d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:6.3f}")


if __name__ == "__main__":
    main()
