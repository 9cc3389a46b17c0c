#!/bin/bash
# Round 2, call 4: fused backward GEMM — parity first (short timeout: a protocol bug traps after ~2 s per wait), then timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q --timeout 120 -x -k "fused" > gpurun_out/r2c4_tests_fused.log 2>&1
rc=$?
tail -n 15 gpurun_out/r2c4_tests_fused.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "fused tests failed (rc $rc): stopping"; exit 0; fi
timeout 300 python tools/bwd_fused_bench.py > gpurun_out/r2c4_fused_bench.log 2>&1
cat gpurun_out/r2c4_fused_bench.log
T="timeout 900"
B="python bench.py --no-cpu-baseline --steps 30"
VITB_BWD_FUSED=0 VITB_WGRAD_STREAM=2 $T $B > gpurun_out/r2c4_b1024_unfused.log 2>&1
VITB_BWD_FUSED=1 VITB_WGRAD_STREAM=2 $T $B --kernel-table gpurun_out/r2c4_ktable_b1024.json > gpurun_out/r2c4_b1024_fused.log 2>&1
VITB_BWD_FUSED=1 VITB_WGRAD_STREAM=0 $T $B > gpurun_out/r2c4_b1024_fused_nostream.log 2>&1
VITB_BWD_FUSED=0 VITB_WGRAD_STREAM=2 $T $B --batch 128 > gpurun_out/r2c4_b128_unfused.log 2>&1
VITB_BWD_FUSED=1 VITB_WGRAD_STREAM=2 $T $B --batch 128 > gpurun_out/r2c4_b128_fused.log 2>&1
VITB_BWD_FUSED=1 VITB_WGRAD_STREAM=2 $T $B --workload t17c100 > gpurun_out/r2c4_t17_fused.log 2>&1
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c4_tests.log 2>&1
for f in gpurun_out/r2c4_b*.log gpurun_out/r2c4_t17*.log gpurun_out/r2c4_tests.log; do echo "== $f"; tail -n 3 $f | cut -c1-330; done
