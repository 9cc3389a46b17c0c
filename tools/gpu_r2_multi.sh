#!/bin/bash
# Round 2, multi-GPU evidence on one 8-GPU box: headline N=2/4/8 (with the data-parallel numeric check), t17c100 N=2/4/8,
# scaled65 N=8 (global batch 4096), per-phase clocks of the fused data-parallel kernel.
mkdir -p gpurun_out
P=29500
run() {  # N out workload...
  N=$1; OUT=$2; shift 2
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --no-cpu-baseline "$@" > gpurun_out/$OUT 2>&1
  echo "== $OUT (rc $?)"; grep '^{' gpurun_out/$OUT | tail -n 1 | cut -c1-250
}
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/r2m2_gpus.txt 2>&1
run 8 r2m2_headline_n8.log
run 8 r2m2_t17c100_n8.log --workload t17c100
run 8 r2m2_scaled65_n8.log --workload scaled65
run 4 r2m2_headline_n4.log
run 4 r2m2_t17c100_n4.log --workload t17c100
run 2 r2m2_headline_n2.log
run 2 r2m2_t17c100_n2.log --workload t17c100
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2m2_headline_n1.log 2>&1; grep '^{' gpurun_out/r2m2_headline_n1.log | tail -n 1 | cut -c1-200
timeout 300 python bench.py --steps 30 --no-cpu-baseline --workload t17c100 > gpurun_out/r2m2_t17c100_n1.log 2>&1; grep '^{' gpurun_out/r2m2_t17c100_n1.log | tail -n 1 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29590 tools/dp_phases.py > gpurun_out/r2m2_dp_phases_n8.log 2>&1
tail -n 12 gpurun_out/r2m2_dp_phases_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 tools/dp_phases.py > gpurun_out/r2m2_dp_phases_n4.log 2>&1
tail -n 8 gpurun_out/r2m2_dp_phases_n4.log
