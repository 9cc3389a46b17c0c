#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -q --timeout 120 -x -k "gemm" > gpurun_out/r2c19_tests_gemm.log 2>&1
rc=$?; tail -n 3 gpurun_out/r2c19_tests_gemm.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|FAILED" gpurun_out/r2c19_tests_gemm.log | head; echo "gemm tests failed"; exit 0; fi
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c19_tests.log 2>&1; tail -n 2 gpurun_out/r2c19_tests.log
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --batch 128 > gpurun_out/r2c19_b128_bn192.log 2>&1
VITB_GEMM_BN192=0 $T $B --batch 128 > gpurun_out/r2c19_b128_bn128.log 2>&1
$T $B --batch 112 > gpurun_out/r2c19_b112_bn192.log 2>&1
VITB_GEMM_BN192=0 $T $B --batch 112 > gpurun_out/r2c19_b112_bn128.log 2>&1
$T $B > gpurun_out/r2c19_b1024.log 2>&1
VITB_GEMM_BN192_STREAM=1 $T $B --kernel-table gpurun_out/r2c19_ktable_stream.json > gpurun_out/r2c19_b1024_stream192.log 2>&1
for f in gpurun_out/r2c19_b*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c19_ktable_stream.json 2>/dev/null | grep -E "gemm_dgrad|graph"
timeout 300 python tools/cublas_shapes.py 8320 2>&1 | grep -v '^{' | head -12
