#!/bin/bash
# ncu full capture (with source counters) of selected kernels from one eager bench step.
mkdir -p gpurun_out
T="timeout 1200"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$T $CMD > gpurun_out/plain.log 2>&1 && \
$T ncu --set full --clock-control none --import-source on -k regex:"${1:-attn_bwd_bf16|attn_fwd_bf16|ln_bwd|gemm_tc_kernel}" -s ${2:-300} -c ${3:-30} -o gpurun_out/prof_sel $CMD > gpurun_out/ncu_sel.log 2>&1
tail -n 3 gpurun_out/ncu_sel.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
