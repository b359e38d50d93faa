#!/usr/bin/env python
"""Per-kernel SASS / resource table of the shipped library: counts of the instructions that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR =
tcgen05.commit, LDGSTS = cp.async, HMMA = legacy mma.sync) plus registers / shared memory / spills from
`cuobjdump -res-usage`.  No GPU needed.

    python tools/sass_table.py [tiny-newsrec_b200/libtinyrec.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDGSTS", "HMMA", "MUFU", "SYNCS", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tiny-newsrec_b200", "libtinyrec.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + ".") or (p == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                counts[cur][p] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = {k: v for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
            fn = None
    names = demangle(list(counts))
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDGSTS", "HMMA", "MUFU"]
    print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels (cuobjdump -sass / -res-usage; static counts, not executed counts)")
    print(f"{'kernel':78s} {'inst':>6s} {'regs':>4s} {'stack':>5s} {'smem':>6s} " + " ".join(f"{c:>7s}" for c in cols))
    tot = collections.Counter()
    for k, c in counts.items():
        u = usage.get(k, {})
        short = re.sub(r"\(.*", "", names.get(k, k))
        short = short.replace("void ", "").replace("tnr::", "")
        print(f"{short[:78]:78s} {c['_total']:6d} {u.get('REG', '?'):>4s} {u.get('STACK', '?'):>5s} {u.get('SHARED', '?'):>6s} "
              + " ".join(f"{c[p]:7d}" for p in cols))
        tot.update(c)
    print(f"{'TOTAL':78s} {tot['_total']:6d} {'':>4s} {'':>5s} {'':>6s} " + " ".join(f"{tot[p]:7d}" for p in cols))


if __name__ == "__main__":
    main()
