#!/bin/bash
# Round 2, call 1: new parity tests, baseline numbers on today's box (headline, B=128, T=17), eager-GPU arm, TS kill criterion.
mkdir -p gpurun_out
T="timeout 600"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
$T python -m pytest tests/test_gpu_round2.py -q --timeout 300 -x > gpurun_out/r2c1_tests.log 2>&1
$T python bench.py --kernel-table gpurun_out/r2c1_ktable_b1024.json > gpurun_out/r2c1_bench.log 2>&1
$T python bench.py --batch 128 --no-cpu-baseline --kernel-table gpurun_out/r2c1_ktable_b128.json > gpurun_out/r2c1_bench_b128.log 2>&1
$T python bench.py --workload t17c100 --batch 128 --no-cpu-baseline > gpurun_out/r2c1_bench_t17_b128.log 2>&1
$T python bench.py --workload t17c100 --no-cpu-baseline --kernel-table gpurun_out/r2c1_ktable_t17.json > gpurun_out/r2c1_bench_t17.log 2>&1
$T python bench.py --impl eager --steps 10 > gpurun_out/r2c1_eager_b1024.log 2>&1
$T python bench.py --impl eager --steps 10 --batch 128 > gpurun_out/r2c1_eager_b128.log 2>&1
VITB_GEMM_TS=1 $T python bench.py --no-cpu-baseline --kernel-table gpurun_out/r2c1_ktable_ts.json > gpurun_out/r2c1_bench_ts.log 2>&1
for f in gpurun_out/r2c1_*.log; do echo "== $f"; tail -n 4 $f | cut -c1-1800; done
