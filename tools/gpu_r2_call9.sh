#!/bin/bash
# Round 2, call 9: persistent attention backward
mkdir -p gpurun_out
T="timeout 900"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -x -k "attention" > gpurun_out/r2c9_tests_attn.log 2>&1
rc=$?; tail -n 5 gpurun_out/r2c9_tests_attn.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "attention tests failed"; exit 0; fi
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c9_tests.log 2>&1; tail -n 3 gpurun_out/r2c9_tests.log
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --kernel-table gpurun_out/r2c9_ktable_b1024.json > gpurun_out/r2c9_b1024.log 2>&1
VITB_ATTN_BWD_ONESHOT=1 $T $B --kernel-table gpurun_out/r2c9_ktable_b1024_oneshot.json > gpurun_out/r2c9_b1024_oneshot.log 2>&1
$T $B --batch 128 > gpurun_out/r2c9_b128.log 2>&1
VITB_ATTN_BWD_ONESHOT=1 $T $B --batch 128 > gpurun_out/r2c9_b128_oneshot.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c9_t17.log 2>&1
VITB_ATTN_BWD_ONESHOT=1 $T $B --workload t17c100 > gpurun_out/r2c9_t17_oneshot.log 2>&1
for f in gpurun_out/r2c9_b*.log gpurun_out/r2c9_t17*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c9_ktable_b1024.json | grep attn
python tools/ktable.py gpurun_out/r2c9_ktable_b1024_oneshot.json | grep attn
