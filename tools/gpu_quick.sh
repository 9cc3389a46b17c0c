#!/bin/bash
# Quick GPU check: GEMM debug, GPU tests, smoke, bench with per-kernel table.
mkdir -p gpurun_out
T="timeout 900"
$T python tools/debug_gemm.py > gpurun_out/debug_gemm.log 2>&1
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/t_gpu.log 2>&1
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "== debug_gemm"; head -30 gpurun_out/debug_gemm.log
for f in gpurun_out/t_gpu.log gpurun_out/smoke.log gpurun_out/bench.log; do echo "== $f"; tail -n 6 $f | cut -c1-1200; done
