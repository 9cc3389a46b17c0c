#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests/test_gpu_round2.py -q --timeout 300 -x -k "all_gradients" > gpurun_out/r2c14_tests_grads.log 2>&1; tail -n 3 gpurun_out/r2c14_tests_grads.log | cut -c1-400
VITB_BWD_FUSED=1 $T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c14_tests_fused_on.log 2>&1; tail -n 2 gpurun_out/r2c14_tests_fused_on.log | cut -c1-300
VITB_PDL=0 VITB_WGRAD_STREAM=0 VITB_DEFER=0 $T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c14_tests_r1_mode.log 2>&1; tail -n 2 gpurun_out/r2c14_tests_r1_mode.log | cut -c1-300
