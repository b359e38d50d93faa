#!/usr/bin/env python
"""Summarise gpurun_out/<tag>/test_report.jsonl (every error the GPU parity tests measured, tests/conftest.py:report)
into a table: per test the worst score / loss error and the worst per-tensor gradient error against its bound.

    python tools/test_report_summary.py gpurun_out/r2b/test_report.jsonl > profiles/r02_test_report_summary.txt
"""
import collections
import json
import sys


def main(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    groups = collections.OrderedDict()
    for r in rows:
        lab = r["label"]
        if ".grad." in lab:
            test, what = lab.split(".grad.")[0], "grad"
            name = lab.split(".grad.")[1]
        else:
            test, _, name = lab.rpartition(".")
            what = "value"
        g = groups.setdefault((test, what), [])
        g.append((r["value"], name, r.get("bound")))
    print(f"{'test':72s} {'kind':6s} {'n':>4s} {'worst':>10s} {'bound':>8s}  where")
    for (test, what), v in groups.items():
        v.sort(key=lambda t: -t[0])
        worst, name, bound = v[0]
        print(f"{test[:72]:72s} {what:6s} {len(v):4d} {worst:10.3e} {bound if bound is not None else float('nan'):8.1e}  {name[-60:]}")


if __name__ == "__main__":
    main(sys.argv[1])
