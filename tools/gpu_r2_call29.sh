#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
B="python bench.py --no-cpu-baseline --steps 40"
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c29_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c29_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["value"])')"; }
B="python bench.py --no-cpu-baseline --steps 40 --batch 128"
run b128_kb1 A=1
run b128_kb8 VITB_WGRAD_MIN_KBLOCKS=8
run b128_kb12 VITB_WGRAD_MIN_KBLOCKS=12
run b128_kb16 VITB_WGRAD_MIN_KBLOCKS=16
run b128_kb24 VITB_WGRAD_MIN_KBLOCKS=24
run b128_kb1b A=1
B="python bench.py --no-cpu-baseline --steps 40 --workload t17c100"
run t17_kb1 A=1
run t17_kb12 VITB_WGRAD_MIN_KBLOCKS=12
run t17_kb24 VITB_WGRAD_MIN_KBLOCKS=24
B="python bench.py --no-cpu-baseline --steps 40 --workload t17c100 --batch 128"
run t17b128_kb1 A=1
run t17b128_kb8 VITB_WGRAD_MIN_KBLOCKS=8
