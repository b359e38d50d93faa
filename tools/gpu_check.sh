#!/bin/bash
# Quick GPU-box visit: parity tests, smoke, the default bench line.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
tag=${1:-check}
o=gpurun_out/$tag
mkdir -p $o
rm -f gpurun_out/test_report.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $o/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=15 > $o/pytest.log 2>&1; echo "pytest exit $?" >> $o/pytest.log; tail -5 $o/pytest.log
cp gpurun_out/test_report.jsonl $o/ 2>/dev/null
timeout 120 python __graft_entry__.py --smoke > $o/smoke.log 2>&1; tail -1 $o/smoke.log
timeout 400 python bench.py > $o/bench_kd4.json 2> $o/bench_kd4.err; tail -c 1500 $o/bench_kd4.json; tail -3 $o/bench_kd4.err
if [ "$2" = "eval" ]; then
  timeout 300 python bench.py --workload eval --no-cpu-baseline > $o/bench_eval.json 2> $o/bench_eval.err; tail -c 1200 $o/bench_eval.json; tail -3 $o/bench_eval.err
  timeout 200 python tools/ue_bench.py > $o/ue_bench.txt 2>&1; cat $o/ue_bench.txt
fi
