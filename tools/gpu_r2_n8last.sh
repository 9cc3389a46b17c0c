#!/bin/bash
# Last 8-GPU lines of round 2 (after the head / loss kernel rewrite and the one-wave tiles): headline, t17c100, scaled65 at N=8, N=1 beside them
mkdir -p gpurun_out
P=29700
run() { N=$1; OUT=$2; shift 2; P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --no-cpu-baseline "$@" > gpurun_out/$OUT 2>&1
  echo "== $OUT (rc $?)"; grep '^{' gpurun_out/$OUT | tail -n 1 | cut -c1-230; grep -o '"dp_check": {[^}]*}' gpurun_out/$OUT | tail -n 1 | cut -c1-300; }
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/r2n8_gpus.txt 2>&1
run 8 r2n8_headline_n8.log
run 8 r2n8_t17c100_n8.log --workload t17c100
run 8 r2n8_scaled65_n8.log --workload scaled65
run 4 r2n8_headline_n4.log
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2n8_headline_n1.log 2>&1; grep '^{' gpurun_out/r2n8_headline_n1.log | tail -n 1 | cut -c1-200
timeout 300 python bench.py --steps 30 --no-cpu-baseline --workload t17c100 > gpurun_out/r2n8_t17c100_n1.log 2>&1; grep '^{' gpurun_out/r2n8_t17c100_n1.log | tail -n 1 | cut -c1-200
