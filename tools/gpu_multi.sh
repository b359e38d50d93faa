#!/bin/bash
# Multi-GPU box visit.  Usage: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh <tag> <N> [tests] [kd4] [kd2] [table] [table_long] [eval]'
tag=$1; n=$2; shift; shift
o=gpurun_out/$tag
mkdir -p $o
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $o/gpu.txt 2>&1
nvidia-smi topo -m > $o/topo.txt 2>&1
suffix=""
run() { # workload, extra args; $suffix distinguishes variants of the same workload
  wl=$1; shift
  f=$o/bench_${wl}_n$n$suffix
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --workload $wl "$@" > $f.json 2> $f.err
  echo "== $wl n=$n$suffix rc=$?"; grep '^{' $f.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['unit'], d['ms_per_step'], 'ms/step; sustained', (d.get('sustained') or {}).get('value'))"; tail -2 $f.err
}
for what in "$@"; do
  case $what in
    tests) timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA > $o/pytest_multi_gpu.log 2>&1; echo "pytest exit $?" >> $o/pytest_multi_gpu.log; tail -8 $o/pytest_multi_gpu.log ;;
    kd4) run kd4 --steps 60 --warmup 5 --no-cpu-baseline ;;
    kd4_1) timeout 400 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-preassembled > $o/bench_kd4_n1.json 2> $o/bench_kd4_n1.err; echo "== kd4 n=1 rc=$?"; grep '^{' $o/bench_kd4_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['unit'], d['ms_per_step'], 'ms/step; sustained', d['sustained']['value'])" ;;
    kd4_nccl_only) suffix=_nccl; TNR_P2P_ALLREDUCE=0 run kd4 --steps 60 --warmup 5 --no-cpu-baseline; suffix="" ;;
    kd4_r2) suffix=_reserve2; TNR_COMM_SM_RESERVE=2 run kd4 --steps 60 --warmup 5 --no-cpu-baseline; suffix="" ;;
    kd4_nccl) NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL run kd4 --steps 5 --warmup 3 --no-cpu-baseline; grep -i "nvls\|channels\|algo" $o/bench_kd4_n$n.err | head -20 > $o/nccl_info.txt ;;
    kd2) run kd2 --steps 60 --warmup 5 --no-cpu-baseline ;;
    table) run table --steps 20 --warmup 3 --no-cpu-baseline ;;
    table_long) run table_long --steps 10 --warmup 3 --no-cpu-baseline ;;
    eval) run eval --steps 40 --warmup 5 --no-cpu-baseline ;;
  esac
done
