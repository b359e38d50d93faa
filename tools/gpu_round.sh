#!/bin/bash
# One GPU-box visit that refreshes every number DESIGN.md / profiles/ quote: parity tests, the default bench line
# (CUDA-graph replay + eager roofline pass + CPU baseline), the reference arm, the secondary workloads, per-op and
# per-kernel bandwidth tables, the ncu launch list of the bench command and ncu --set full of the head kernels.
# Usage (repo root): gpurun --timeout 1800 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-round}
o=gpurun_out/$tag
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $o/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $o/pytest.log 2>&1; echo "pytest exit $?" >> $o/pytest.log; tail -2 $o/pytest.log
timeout 120 python __graft_entry__.py --smoke > $o/smoke.log 2>&1; tail -1 $o/smoke.log
timeout 400 python bench.py > $o/bench_kd4.json 2> $o/bench_kd4.err; tail -c 300 $o/bench_kd4.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_ref.json 2> $o/bench_ref.err
timeout 300 python bench.py --workload kd2 --no-cpu-baseline > $o/bench_kd2.json 2> $o/bench_kd2.err
timeout 300 python bench.py --workload eval > $o/bench_eval.json 2> $o/bench_eval.err
timeout 300 python bench.py --workload table > $o/bench_table.json 2> $o/bench_table.err
timeout 300 python bench.py --workload table_long > $o/bench_table_long.json 2> $o/bench_table_long.err
timeout 200 python tools/step_profile.py --out $o/step_profile.json > $o/step_profile.txt 2>&1
timeout 200 python tools/kernel_bw.py > $o/kernel_bw.txt 2>&1; cp gpurun_out/kernel_bw.json $o/ 2>/dev/null
timeout 200 python tools/ue_bench.py > $o/ue_bench.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/ncu_bench.log 2>&1
python tools/ncu_summary.py $o/launches.csv > $o/launches_summary.txt 2>&1
bash tools/ncu_kernels.sh ${tag}_head "ue_logits|ue_pool|eval_metrics" 6 3 python bench.py --workload eval --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_raw_pick.py gpurun_out/${tag}_head/raw.csv > $o/ncu_head_kernels.txt 2>&1
for f in bench_kd2 bench_eval bench_table bench_table_long; do python - <<P
import json
try:
    d = json.load(open("$o/$f.json")); print("$f", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$f failed", e)
P
done
