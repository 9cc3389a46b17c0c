#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_round2.py -q --timeout 100 -x -k "wgrad or all_gradients" > gpurun_out/r2c37_tests.log 2>&1; tail -n 1 gpurun_out/r2c37_tests.log | cut -c1-150
timeout 200 python bench.py --kernel-table gpurun_out/r2c37_ktable_b1024.json > gpurun_out/r2c37_bench.log 2>&1
grep '^{' gpurun_out/r2c37_bench.log | tail -n 1 | cut -c1-220
