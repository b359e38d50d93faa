#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel (no GPU needed): compiles the .cu with -lineinfo to a
cubin, disassembles it with `nvdisasm -g` and attributes every instruction to the innermost source line.  This is how the
tile-addressing bloat of the attention kernels was found (a quarter of the backward kernel's instructions were index
arithmetic of 56 copies).  Usage: python tools/sass_lines.py tiny-newsrec_b200/csrc/attention.cu attn_bwd_kernel [N]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    src, kern = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    extra = os.environ.get("TNR_EXTRA_NVCC_FLAGS", "").split()
    with tempfile.TemporaryDirectory() as d:
        cubin = os.path.join(d, "k.cubin")
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-cubin"] + extra +
                       ["-o", cubin, src], check=True)
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], check=True, capture_output=True, text=True).stdout
    lines = dis.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l]
    if not starts:
        print("no kernel matches; have:", sorted({l.split(",")[0][6:] for l in lines if l.startswith("\t.section\t.text.")})[:40])
        return
    for st in starts:
        end = next((i for i in range(st + 1, len(lines)) if lines[i].startswith("//---------------------")), len(lines))
        cnt = collections.Counter()
        cur = None
        for line in lines[st:end]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
                cnt[cur] += 1
        print(lines[st][:150])
        print("instructions:", sum(cnt.values()))
        cache = {}
        for (f, l), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:topn]:
            path = os.path.join(os.path.dirname(src), f)
            if path not in cache:
                cache[path] = open(path).read().split("\n") if os.path.exists(path) else []
            text = cache[path][l - 1].strip()[:100] if l <= len(cache[path]) else ""
            print(f"{c:6d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
