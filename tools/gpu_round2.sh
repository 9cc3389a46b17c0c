#!/bin/bash
# Full GPU check: tests, smoke, bench, ncu launch list and a full capture of the heaviest kernels.
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/t_gpu.log 2>&1
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json > gpurun_out/bench.log 2>&1
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$T $CMD > gpurun_out/plain.log 2>&1 && \
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$T $CMD > gpurun_out/plain2.log 2>&1 && \
$T ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_bf16|attn_bwd_bf16|ln_bwd|rows_colsum" -s 60 -c 40 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
for f in gpurun_out/t_gpu.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/ncu_list.log gpurun_out/ncu_full.log; do echo "== $f"; tail -n 5 $f | cut -c1-1500; done
ls -la gpurun_out
