#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
B="python bench.py --no-cpu-baseline --steps 40"
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c24_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c24_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["clocks"]["sm_mhz"])')"; }
run base A=1
run pf0 VITB_GEMM_PF_TILES=0
run pf1 VITB_GEMM_PF_TILES=1
run pf3 VITB_GEMM_PF_TILES=3
run pf4 VITB_GEMM_PF_TILES=4
run pf6 VITB_GEMM_PF_TILES=6
run pfin VITB_GEMM_PF_IN=1
run pfin4 VITB_GEMM_PF_IN=1 VITB_GEMM_PF_TILES=4
run kb4 VITB_GEMM_PF_KBLOCKS=4
run kb16 VITB_GEMM_PF_KBLOCKS=16
run kb0 VITB_GEMM_PF_KBLOCKS=0
run base2 A=1
