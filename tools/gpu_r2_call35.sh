#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c35_tests.log 2>&1; tail -n 2 gpurun_out/r2c35_tests.log | cut -c1-200
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c35_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c35_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["value"])')"; }
B="python bench.py --no-cpu-baseline --steps 30"
run b1024_rule A=1
run b1024_old VITB_WGRAD_MAX_SPLITS=-1
B="python bench.py --no-cpu-baseline --steps 30 --batch 512"
run b512_rule A=1
