#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and mean time, share."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:110]


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = [r for r in load(path) if lo <= r[0] < hi]
    agg = OrderedDict()
    for _, name, us, grid, block in rows:
        key = short(name)
        a = agg.setdefault(key, [0, 0.0, set()])
        a[0] += 1
        a[1] += us
        a[2].add(grid)
    total = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {total:.1f} us in total (IDs {lo}..{hi})")
    for k, (c, us, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f} us {100 * us / total:5.1f}%  n={c:5d}  mean {us / c:8.2f} us  {k}  grids={sorted(grids)[:4]}")


if __name__ == "__main__":
    main()
