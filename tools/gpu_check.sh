#!/bin/bash
# One GPU-box visit: parity tests, bench, per-op profile, ncu launch list + full capture of the hot kernels.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json
timeout 200 python tools/step_profile.py --out $out/step_profile.json > $out/step_profile.txt 2>&1; head -40 $out/step_profile.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_bench.log 2>&1
python tools/ncu_summary.py $out/launches.csv > $out/launches_summary.txt 2>&1
if [ -n "$NCU_FULL" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$NCU_FULL" -s 200 -c 24 -o $out/full \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
fi
