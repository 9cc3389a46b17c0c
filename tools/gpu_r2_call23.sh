#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c23_tests.log 2>&1; tail -n 3 gpurun_out/r2c23_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c23_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
NT="env VITB_LIB_PATH=$PWD/gpurun_in_libvitb200_notrim.so"
for rep in 1 2; do
$T $B --batch 128 > gpurun_out/r2c23_b128_trim_$rep.log 2>&1
$NT $T $B --batch 128 > gpurun_out/r2c23_b128_notrim_$rep.log 2>&1
done
$T $B > gpurun_out/r2c23_b1024_trim.log 2>&1
$NT $T $B > gpurun_out/r2c23_b1024_notrim.log 2>&1
$T $B --workload t17c100 --batch 128 > gpurun_out/r2c23_t17b128_trim.log 2>&1
$NT $T $B --workload t17c100 --batch 128 > gpurun_out/r2c23_t17b128_notrim.log 2>&1
for f in gpurun_out/r2c23_*.log; do case $f in *tests*) continue;; esac; echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
timeout 300 python tools/cublas_shapes.py 8320 2>&1 | grep -v '^{' | head -12
