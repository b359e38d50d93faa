#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time,
launch count and share.  Usage: python tools/ncu_summary.py gpurun_out/launches.csv [header text]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit.startswith("n") else v * 1e3 if unit.startswith("m") else v
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, us))
    agg = collections.OrderedDict()
    for n, us in rows:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += us
    tot = sum(v[1] for v in agg.values())
    if len(sys.argv) > 2:
        print("# " + " ".join(sys.argv[2:]))
    print(f"total {tot:.1f} us over {len(rows)} launches")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f} us {c:5d}x {100 * us / tot:5.1f}%  {us / c:8.1f} us/launch  {n[:90]}")


if __name__ == "__main__":
    main()
