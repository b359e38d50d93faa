#!/usr/bin/env python
"""Summarise tools/ncu_step.sh output: one KD train step (from one embed_ln launch to the next), per kernel name the
launch count, total time, DRAM bytes read + written (ncu dram__bytes_*), achieved DRAM GB/s and its fraction of the
measured HBM peak (MEASURED_PEAKS.json), tensor-pipe activity.  Usage: python tools/ncu_step_summary.py step_kernels.csv"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        key = (int(r[ix["ID"]]), r[ix["Kernel Name"]])
        d = per.setdefault(key, {})
        v = r[ix["Metric Value"]].replace(",", "")
        unit = r[ix["Metric Unit"]]
        try:
            v = float(v)
        except ValueError:
            continue
        name = r[ix["Metric Name"]]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if name.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        d[name] = v
    launches = [(k[0], k[1], d) for k, d in per.items()]
    starts = [i for i, (_, n, _) in enumerate(launches) if "embed_ln" in n]
    if len(starts) >= 2:
        launches = launches[starts[0]:starts[1]]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        peak = 6540.0
    agg = collections.OrderedDict()
    for _, n, d in launches:
        short = re.sub(r"\(.*", "", n).replace("void ", "").replace("tnr::", "")
        short = re.sub(r"at::native::.*?(\w+_kernel\w*|\w+Functor\w*).*", r"aten::\1", short)[:70]
        a = agg.setdefault(short, dict(n=0, us=0.0, rd=0.0, wr=0.0, tp=0.0))
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print(f"one KD train step under ncu (--cache-control none, --clock-control none): {len(launches)} launches, {tot:.0f} us; HBM peak {peak:.0f} GB/s")
    print(f"{'kernel':70s} {'n':>3s} {'us':>8s} {'%':>5s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'of HBM':>6s} {'tensor%':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
        print(f"{k:70s} {a['n']:3d} {a['us']:8.1f} {100 * a['us'] / tot:5.1f} {a['rd'] / 1e6:8.1f} {a['wr'] / 1e6:8.1f} {gbs:7.0f} {gbs / peak:6.2f} "
              f"{a['tp'] / a['us'] if a['us'] > 0 else 0:7.1f}")


if __name__ == "__main__":
    main()
