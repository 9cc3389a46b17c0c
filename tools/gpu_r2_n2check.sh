#!/bin/bash
# 2-GPU check after the last code changes of round 2: the two-rank test of the GPU suite, and bench lines (with dp_check) at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q --timeout 300 -k "two_rank" > gpurun_out/r2n2_tests.log 2>&1; tail -n 2 gpurun_out/r2n2_tests.log
P=29600
run() { OUT=$1; shift; P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 30 --no-cpu-baseline "$@" > gpurun_out/$OUT 2>&1
  echo "== $OUT (rc $?)"; grep '^{' gpurun_out/$OUT | tail -n 1 | cut -c1-230; grep -o '"dp_check": {[^}]*}' gpurun_out/$OUT | tail -n 1 | cut -c1-300; }
run r2n2_headline_n2.log
run r2n2_t17c100_n2.log --workload t17c100
run r2n2_headline_n2_drop.log --dropout 0.1
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2n2_headline_n1.log 2>&1; grep '^{' gpurun_out/r2n2_headline_n1.log | tail -n 1 | cut -c1-200
