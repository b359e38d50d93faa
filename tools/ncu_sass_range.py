#!/usr/bin/env python
"""Stall-reason totals over a range of SASS lines of one kernel from an `ncu --page source --csv --print-source sass`
export.  Usage: python tools/ncu_sass_range.py source_sass.csv.gz "<kernel substring>" <first> <last> [occurrence]"""
import collections
import csv
import gzip
import io
import sys


def main():
    path, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    occ = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    txt = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
    lines = txt.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
    blocks = [(lines[starts[k]], lines[starts[k] + 1:starts[k + 1]]) for k in range(len(starts) - 1)]
    name, body = [b for b in blocks if pat in b[0]][occ]
    rd = list(csv.reader(io.StringIO("\n".join(body))))
    hdr = rd[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    ops = collections.Counter()
    samples = total = executed = 0
    for i, r in enumerate(rd[1:]):
        if len(r) < len(hdr):
            continue
        s = int(r[ix["# Samples"]] or 0)
        total += s
        if lo <= i <= hi:
            samples += s
            executed += int(r[ix["Warp Instructions Executed"]] or 0) if "Warp Instructions Executed" in ix else 0
            ops[r[ix["Source"]].split()[1 if r[ix["Source"]].lstrip().startswith("@") else 0].split(".")[0]] += 1
            for c in stall_cols:
                if r[ix[c]] not in ("", "0"):
                    tot[c[6:]] += int(r[ix[c]])
    print(f"lines {lo}-{hi}: {samples} of {total} samples, {executed} warp instructions executed")
    print("stalls:", ", ".join(f"{k} {v} ({100.0 * v / max(samples, 1):.0f}%)" for k, v in tot.most_common(10)))
    print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main()
