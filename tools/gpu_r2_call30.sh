#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c30_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c30_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["value"])')"; }
B="python bench.py --no-cpu-baseline --steps 30"
run b1024_s24 A=1
run b1024_s16 VITB_WGRAD_MAX_SPLITS=16
run b1024_s12 VITB_WGRAD_MAX_SPLITS=12
run b1024_s8 VITB_WGRAD_MAX_SPLITS=8
B="python bench.py --no-cpu-baseline --steps 30 --batch 128"
run b128_s12 VITB_WGRAD_MAX_SPLITS=12
run b128_s10 VITB_WGRAD_MAX_SPLITS=10
run b128_s8 VITB_WGRAD_MAX_SPLITS=8
run b128_s6 VITB_WGRAD_MAX_SPLITS=6
B="python bench.py --no-cpu-baseline --steps 30 --workload t17c100"
run t17_s16 VITB_WGRAD_MAX_SPLITS=16
run t17_s11 VITB_WGRAD_MAX_SPLITS=11
run t17_s8 VITB_WGRAD_MAX_SPLITS=8
run t17_s6 VITB_WGRAD_MAX_SPLITS=6
B="python bench.py --no-cpu-baseline --steps 30 --workload t17c100 --batch 128"
run t17b128_s8 VITB_WGRAD_MAX_SPLITS=8
run t17b128_s4 VITB_WGRAD_MAX_SPLITS=4
B="python bench.py --no-cpu-baseline --steps 20 --workload scaled65"
run sc65_s24 A=1
run sc65_s8 VITB_WGRAD_MAX_SPLITS=8
