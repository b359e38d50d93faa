#!/bin/bash
# ncu --set full of ONE launch of a kernel, exported as CSV (raw metrics + SASS source page) on the box.
# Usage: gpurun -- 'bash tools/ncu_kernel.sh <tag> <kernel regex> <skip> <python command...>'
tag=$1; shift; kre=$1; shift; skip=$1; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"$kre" -s $skip -c 1 -o /tmp/one_kernel "$@" > $out/ncu.log 2>&1
ncu -i /tmp/one_kernel.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
ncu -i /tmp/one_kernel.ncu-rep --page details --csv > $out/details.csv 2>/dev/null
ncu -i /tmp/one_kernel.ncu-rep --page source --csv --print-source sass | gzip > $out/source_sass.csv.gz
ls -la $out
