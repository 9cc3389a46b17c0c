#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c12_tests.log 2>&1; tail -n 2 gpurun_out/r2c12_tests.log
B="python bench.py --no-cpu-baseline --steps 20"
$T $B --workload scaled17 > gpurun_out/r2c12_s17.log 2>&1
VITB_DEFER=0 $T $B --workload scaled17 > gpurun_out/r2c12_s17_nodefer.log 2>&1
$T $B --workload scaled65 > gpurun_out/r2c12_s65.log 2>&1
VITB_DEFER=0 $T $B --workload scaled65 > gpurun_out/r2c12_s65_nodefer.log 2>&1
$T $B > gpurun_out/r2c12_b1024.log 2>&1
VITB_DEFER=0 $T $B > gpurun_out/r2c12_b1024_nodefer.log 2>&1
$T $B --batch 128 > gpurun_out/r2c12_b128.log 2>&1
VITB_DEFER=0 $T $B --batch 128 > gpurun_out/r2c12_b128_nodefer.log 2>&1
for f in gpurun_out/r2c12_*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
