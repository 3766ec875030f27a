#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch
count, total and share of device time.  Usage: summarize_launches.py launches.csv [first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "")
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    if limit:
        rows = rows[:limit]
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, ns, ns])
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
        a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms device time (cold-cache, serialised)")
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'min_us':>9s} {'max_us':>9s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:58]:58s} {a[0]:8d} {a[1] / 1e3:10.1f} {100 * a[1] / total:6.1f}% {a[2] / 1e3:9.1f} {a[3] / 1e3:9.1f}")


if __name__ == "__main__":
    main()
