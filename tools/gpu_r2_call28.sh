#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -q --timeout 120 -x -k "gemm" > gpurun_out/r2c28_tests_gemm.log 2>&1
rc=$?; tail -n 3 gpurun_out/r2c28_tests_gemm.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|FAILED" gpurun_out/r2c28_tests_gemm.log | head; echo "gemm tests failed"; exit 0; fi
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c28_tests.log 2>&1; tail -n 3 gpurun_out/r2c28_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c28_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
OLD="env VITB_LIB_PATH=$PWD/gpurun_in_lib_before_wbars.so"
for rep in 1 2; do
$T $B > gpurun_out/r2c28_b1024_new_$rep.log 2>&1
$OLD $T $B > gpurun_out/r2c28_b1024_old_$rep.log 2>&1
done
$T $B --workload t17c100 > gpurun_out/r2c28_t17_new.log 2>&1
$OLD $T $B --workload t17c100 > gpurun_out/r2c28_t17_old.log 2>&1
$T $B --workload scaled65 > gpurun_out/r2c28_scaled65_new.log 2>&1
$OLD $T $B --workload scaled65 > gpurun_out/r2c28_scaled65_old.log 2>&1
for f in gpurun_out/r2c28_*new*.log gpurun_out/r2c28_*old*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
