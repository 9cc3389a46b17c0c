#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
run() { name=$1; shift; env "$@" $T $B > gpurun_out/r2c32_$name.log 2>&1; echo "$name $(grep '^{' gpurun_out/r2c32_$name.log | tail -n 1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["step_ms"]["p10"], d["step_ms"]["p50"], d["value"])')"; }
B="python bench.py --no-cpu-baseline --steps 40 --batch 128"
run b128_base A=1
run b128_wide9 VITB_WGRAD_WIDE_MAX_ROWBLOCKS=9
run b128_noflushpl VITB_FLUSH_PER_LAYER=0
run b128_nooptpl VITB_OPT_PER_LAYER=0
run b128_lngelu VITB_LN_GELU_FUSED=1
run b128_bn192off VITB_GEMM_BN192=0
run b128_base2 A=1
B="python bench.py --no-cpu-baseline --steps 40 --workload t17c100 --batch 128"
run t17b128_base A=1
run t17b128_wide9 VITB_WGRAD_WIDE_MAX_ROWBLOCKS=9
run t17b128_lngelu VITB_LN_GELU_FUSED=1
B="python bench.py --no-cpu-baseline --steps 40 --workload t17c100"
run t17_wide9 VITB_WGRAD_WIDE_MAX_ROWBLOCKS=9
run t17_lngelu VITB_LN_GELU_FUSED=1
run t17_base A=1
