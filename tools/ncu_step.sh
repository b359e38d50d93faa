#!/bin/bash
# Per-kernel time + DRAM traffic of one eagerly launched KD train step (warm caches: --cache-control none), every kernel.
# Usage: gpurun -- 'bash tools/ncu_step.sh <tag>'   -> gpurun_out/<tag>/step_kernels.csv + step_kernels.txt
tag=${1:-step}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none --cache-control none -s 500 -c 300 --csv --log-file $out/step_kernels.csv \
  python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-graph --no-preassembled > $out/ncu_step.log 2>&1
python tools/ncu_step_summary.py $out/step_kernels.csv > $out/step_kernels.txt 2>&1
tail -45 $out/step_kernels.txt
