#!/bin/bash
# ncu --set full of the first N launches matching a kernel regex (after <skip> matching launches), CSV exports on the box.
# Usage: gpurun -- 'bash tools/ncu_kernels.sh <tag> <kernel regex> <skip> <count> <python command...>'
tag=$1; shift; kre=$1; shift; skip=$1; shift; cnt=$1; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -o /tmp/kernels "$@" > $out/ncu.log 2>&1
ncu -i /tmp/kernels.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
ncu -i /tmp/kernels.ncu-rep --page source --csv --print-source sass | gzip > $out/source_sass.csv.gz
ls -la $out
