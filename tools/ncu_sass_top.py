#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an `ncu --page source --csv --print-source sass` export
(optionally gzipped).  Usage: python tools/ncu_sass_top.py source_sass.csv.gz "<kernel name substring>" [N] [occurrence]"""
import csv
import gzip
import io
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    txt = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
    lines = txt.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
    blocks = [(lines[starts[k]], lines[starts[k] + 1:starts[k + 1]]) for k in range(len(starts) - 1)]
    sel = [b for b in blocks if pat in b[0]]
    if not sel:
        print("no kernel matches; have:", sorted({b[0][:140] for b in blocks}))
        return
    name, body = sel[occ]
    rd = list(csv.reader(io.StringIO("\n".join(body))))
    hdr = rd[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    rows = []
    for r in rd[1:]:
        if len(r) < len(hdr):
            continue
        try:
            s = int(r[ix["# Samples"]])
        except ValueError:
            continue
        rows.append((s, r))
    total = sum(s for s, _ in rows)
    print(name[:160])
    print("total samples", total, "instructions", len(rows))
    order = sorted(range(len(rows)), key=lambda i: -rows[i][0])[:topn]
    for i in sorted(order):
        s, r = rows[i]
        st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]] not in ("", "0")), reverse=True)[:3]
        print(f"{i:5d} {100.0 * s / max(total, 1):5.1f}%  {r[ix['Source']][:70]:70s} {st}")


if __name__ == "__main__":
    main()
