#!/bin/bash
# Round-end evidence on one B200.  One profiler per call:
#   bash tools/gpu_final.sh          tests, smoke, bench (+ per-kernel table), reference arm, ncu launch list
#   bash tools/gpu_final.sh full     ncu --set full capture of the hot kernels (backward window, then nothing else)
mkdir -p gpurun_out
T="timeout 900"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
if [ "$1" = "full" ]; then
  $T $CMD > gpurun_out/plain2.log 2>&1 && \
  # the report goes to /tmp: gpurun only copies back 64 MiB, so export the per-kernel metrics as CSV and keep the report only if small
  $T ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_bf16|attn_bwd_bf16|ln_bwd|gelu_bwd|ln_fwd|adam" -s ${2:-330} -c ${3:-40} -o /tmp/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
  tail -n 3 gpurun_out/ncu_full.log | cut -c1-300
  ncu -i /tmp/prof_r1.ncu-rep --page raw --csv > gpurun_out/ncu_full_raw.csv 2> gpurun_out/ncu_export.log
  ls -la /tmp/prof_r1.ncu-rep gpurun_out/ncu_full_raw.csv
  [ $(stat -c %s /tmp/prof_r1.ncu-rep) -lt 40000000 ] && cp /tmp/prof_r1.ncu-rep gpurun_out/
  exit 0
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/t_gpu.log 2>&1
$T python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json > gpurun_out/bench.log 2>&1
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
$T $CMD > gpurun_out/plain.log 2>&1 && \
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
for f in gpurun_out/t_gpu.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/ncu_list.log; do echo "== $f"; tail -n 3 $f | cut -c1-700; done
head -n 12 gpurun_out/launch_summary.txt
