#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -x -k "patch or gemm_fwd" > gpurun_out/r2c13_tests_patch.log 2>&1
rc=$?; tail -n 3 gpurun_out/r2c13_tests_patch.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/r2c13_tests_patch.log | head; echo "patch tests failed"; exit 0; fi
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c13_tests.log 2>&1; tail -n 2 gpurun_out/r2c13_tests.log
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --kernel-table gpurun_out/r2c13_ktable_b1024.json > gpurun_out/r2c13_b1024.log 2>&1
$T $B --batch 128 > gpurun_out/r2c13_b128.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c13_t17.log 2>&1
for f in gpurun_out/r2c13_b*.log gpurun_out/r2c13_t17.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c13_ktable_b1024.json 2>/dev/null | grep -E "patch|gelu|graph"
