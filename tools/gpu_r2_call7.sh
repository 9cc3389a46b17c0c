#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q --timeout 300 -x -k "guard" > gpurun_out/r2c7_guard.log 2>&1; tail -n 3 gpurun_out/r2c7_guard.log
timeout 600 python tools/cublas_shapes.py > gpurun_out/r2c7_cublas_b1024.log 2>&1; cat gpurun_out/r2c7_cublas_b1024.log | cut -c1-200 | head -20
timeout 600 python tools/cublas_shapes.py 8320 > gpurun_out/r2c7_cublas_b128.log 2>&1; cat gpurun_out/r2c7_cublas_b128.log | cut -c1-200 | head -20
