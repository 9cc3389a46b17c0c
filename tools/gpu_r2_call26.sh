#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -x -k "ls_ce or head" > gpurun_out/r2c26_tests_head.log 2>&1
tail -n 3 gpurun_out/r2c26_tests_head.log | cut -c1-300
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c26_tests.log 2>&1; tail -n 3 gpurun_out/r2c26_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c26_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --workload t17c100 --kernel-table gpurun_out/r2c26_ktable_t17.json > gpurun_out/r2c26_t17.log 2>&1
$T $B --workload t17c100 --batch 128 > gpurun_out/r2c26_t17_b128.log 2>&1
$T $B --batch 128 > gpurun_out/r2c26_b128.log 2>&1
$T $B > gpurun_out/r2c26_b1024.log 2>&1
for f in gpurun_out/r2c26_t17*.log gpurun_out/r2c26_b*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c26_ktable_t17.json 2>/dev/null | grep -E "ls_ce|out_f32|dy_f32|graph|sum eager"
