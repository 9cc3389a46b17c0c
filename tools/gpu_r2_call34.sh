#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c34_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c34_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["value"])')"; }
B="python bench.py --no-cpu-baseline --steps 30 --batch 256"
run b256_rule A=1
run b256_old VITB_WGRAD_MAX_SPLITS=-1
run b256_s8 VITB_WGRAD_MAX_SPLITS=8
B="python bench.py --no-cpu-baseline --steps 30 --batch 512"
run b512_rule A=1
run b512_old VITB_WGRAD_MAX_SPLITS=-1
run b512_s12 VITB_WGRAD_MAX_SPLITS=12
