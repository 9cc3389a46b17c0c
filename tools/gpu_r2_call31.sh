#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c31_tests.log 2>&1; tail -n 3 gpurun_out/r2c31_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c31_tests.log | head
$T python __graft_entry__.py smoke > gpurun_out/r2c31_smoke.log 2>&1; tail -n 1 gpurun_out/r2c31_smoke.log
$T python bench.py --kernel-table gpurun_out/r2c31_ktable_b1024.json > gpurun_out/r2c31_bench.log 2>&1
$T python bench.py --batch 128 --no-cpu-baseline --kernel-table gpurun_out/r2c31_ktable_b128.json > gpurun_out/r2c31_bench_b128.log 2>&1
$T python bench.py --workload t17c100 --no-cpu-baseline > gpurun_out/r2c31_bench_t17c100.log 2>&1
$T python bench.py --workload t17c100 --batch 128 --no-cpu-baseline > gpurun_out/r2c31_bench_t17c100_b128.log 2>&1
$T python bench.py --workload scaled17 --no-cpu-baseline > gpurun_out/r2c31_bench_scaled17.log 2>&1
for f in gpurun_out/r2c31_bench*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-220; done
