#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c33_tests.log 2>&1; tail -n 3 gpurun_out/r2c33_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c33_tests.log | head
$T python __graft_entry__.py smoke > gpurun_out/r2c33_smoke.log 2>&1; tail -n 1 gpurun_out/r2c33_smoke.log
B="python bench.py --no-cpu-baseline --steps 40"
$T $B --batch 128 > gpurun_out/r2c33_bench_b128.log 2>&1
$T $B --workload t17c100 --batch 128 > gpurun_out/r2c33_bench_t17c100_b128.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c33_bench_t17c100.log 2>&1
$T $B > gpurun_out/r2c33_bench.log 2>&1
for f in gpurun_out/r2c33_bench*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-220; done
grep -o '"gpu_launches": [0-9]*' gpurun_out/r2c33_bench.log | tail -1
