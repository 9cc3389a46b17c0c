#!/bin/bash
# First GPU trip: kernel unit tests (GEMM variants isolated per process), parity tests, smoke, bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout 600"
$T python tools/debug_gemm.py > gpurun_out/debug_gemm.log 2>&1
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not gemm" --timeout 120 > gpurun_out/t_kernels_other.log 2>&1
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_fwd" --timeout 120 > gpurun_out/t_gemm_fwd.log 2>&1
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_dgrad" --timeout 120 > gpurun_out/t_gemm_dgrad.log 2>&1
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_wgrad or gemm_head" --timeout 120 > gpurun_out/t_gemm_wgrad.log 2>&1
$T python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32" --timeout 300 > gpurun_out/t_parity_fp32.log 2>&1
$T python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not fp32" --timeout 300 > gpurun_out/t_parity_bf16.log 2>&1
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
$T python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernels_b1024.json > gpurun_out/bench.log 2>&1
for f in gpurun_out/t_*.log gpurun_out/smoke.log gpurun_out/bench.log; do echo "== $f"; tail -n 4 $f; done
echo "== debug_gemm"; cat gpurun_out/debug_gemm.log | head -60
