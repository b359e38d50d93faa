#!/bin/bash
# ncu --set full over the GEMM launches of one train step (skip 3 warm-up steps), exported as CSV on the box
# (the .ncu-rep itself is too large to bring back).  Usage: gpurun -- 'bash tools/ncu_gemm.sh tag'
tag=${1:-gemm}
out=gpurun_out/$tag
mkdir -p $out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 111 -c 37 -o /tmp/gemm_step \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $out/ncu.log 2>&1
ncu -i /tmp/gemm_step.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
ncu -i /tmp/gemm_step.ncu-rep --page source --csv --print-source sass > $out/source_sass.csv 2>/dev/null
python - "$out" <<'PY'
import sys, os, gzip, shutil
out = sys.argv[1]
p = os.path.join(out, "source_sass.csv")
with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as g:
    shutil.copyfileobj(f, g)
os.remove(p)
print("sizes", {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})
PY
