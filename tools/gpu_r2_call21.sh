#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c21_tests.log 2>&1; tail -n 3 gpurun_out/r2c21_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c21_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --kernel-table gpurun_out/r2c21_ktable.json > gpurun_out/r2c21_b1024_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B > gpurun_out/r2c21_b1024_unfused.log 2>&1
$T $B --batch 128 --kernel-table gpurun_out/r2c21_ktable_b128.json > gpurun_out/r2c21_b128_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B --batch 128 --kernel-table gpurun_out/r2c21_ktable_b128_unfused.json > gpurun_out/r2c21_b128_unfused.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c21_t17_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B --workload t17c100 > gpurun_out/r2c21_t17_unfused.log 2>&1
for f in gpurun_out/r2c21_b*.log gpurun_out/r2c21_t17*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-230; done
for k in r2c21_ktable r2c21_ktable_b128 r2c21_ktable_b128_unfused; do echo "-- $k"; python tools/ktable.py gpurun_out/$k.json 2>/dev/null | grep -E "layernorm_bwd|gelu_bwd|graph" | head; done
