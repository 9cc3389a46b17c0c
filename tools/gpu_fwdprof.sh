#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-1500
python tools/ktable.py 2>/dev/null | head -24
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$T $CMD > gpurun_out/plain2.log 2>&1 && \
$T ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_bf16|attn_bwd_bf16|ln_bwd|gelu_bwd|ln_fwd|adam" -s 416 -c 15 -o gpurun_out/prof_r1_fwd $CMD > gpurun_out/ncu_full_fwd.log 2>&1
tail -n 2 gpurun_out/ncu_full_fwd.log | cut -c1-200
