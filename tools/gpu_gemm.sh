#!/bin/bash
# GEMM-only GPU check: correctness probe, GEMM unit tests, timeline, then the bench with the per-kernel table.
mkdir -p gpurun_out
T="timeout 600"
$T python tools/debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "debug_gemm rc=$?"; cat gpurun_out/debug_gemm.log | cut -c1-150
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "gemm or patch" > gpurun_out/t_gemm.log 2>&1; tail -n 3 gpurun_out/t_gemm.log
$T python tools/gemm_timeline.py > gpurun_out/timeline.log 2>&1; cat gpurun_out/timeline.log | cut -c1-260
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 2 gpurun_out/bench.log | cut -c1-400
