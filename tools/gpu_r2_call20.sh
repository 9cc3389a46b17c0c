#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
timeout 600 python -m pytest tests/test_gpu_round2.py -q --timeout 300 -x -k "dropout or second_output or chains_agree" > gpurun_out/r2c20_tests_new.log 2>&1
rc=$?; tail -n 3 gpurun_out/r2c20_tests_new.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|FAILED" gpurun_out/r2c20_tests_new.log | head -20; fi
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2c20_tests.log 2>&1; tail -n 4 gpurun_out/r2c20_tests.log | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/r2c20_tests.log | head
B="python bench.py --no-cpu-baseline --steps 30"
$T $B > gpurun_out/r2c20_b1024_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B > gpurun_out/r2c20_b1024_unfused.log 2>&1
$T $B --batch 128 > gpurun_out/r2c20_b128_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B --batch 128 > gpurun_out/r2c20_b128_unfused.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c20_t17_fused.log 2>&1
VITB_LN_GELU_FUSED=0 $T $B --workload t17c100 > gpurun_out/r2c20_t17_unfused.log 2>&1
$T $B --kernel-table gpurun_out/r2c20_ktable.json > gpurun_out/r2c20_b1024_kt.log 2>&1
for f in gpurun_out/r2c20_b*.log gpurun_out/r2c20_t17*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-230; done
python tools/ktable.py gpurun_out/r2c20_ktable.json 2>/dev/null | grep -E "ln_bwd|layernorm_bwd|gelu_bwd|graph" | head
