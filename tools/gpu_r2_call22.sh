#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c22_tests.log 2>&1; tail -n 3 gpurun_out/r2c22_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c22_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --dropout 0.1 > gpurun_out/r2c22_drop_fused.log 2>&1
VITB_DROP_FUSED=0 $T $B --dropout 0.1 > gpurun_out/r2c22_drop_unfused.log 2>&1
$T $B --dropout 0.1 --batch 128 > gpurun_out/r2c22_drop_b128_fused.log 2>&1
VITB_DROP_FUSED=0 $T $B --dropout 0.1 --batch 128 > gpurun_out/r2c22_drop_b128_unfused.log 2>&1
$T $B > gpurun_out/r2c22_b1024.log 2>&1
$T $B --batch 128 > gpurun_out/r2c22_b128.log 2>&1
$T $B --dropout 0.1 --kernel-table gpurun_out/r2c22_ktable_drop.json > gpurun_out/r2c22_drop_kt.log 2>&1
for f in gpurun_out/r2c22_*.log; do case $f in *tests*) continue;; esac; echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-230; done
python tools/ktable.py gpurun_out/r2c22_ktable_drop.json 2>/dev/null | head -30
