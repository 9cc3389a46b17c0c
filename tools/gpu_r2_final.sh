#!/bin/bash
# Round-2 evidence on one B200.  One profiler per call:
#   bash tools/gpu_r2_final.sh          tests, smoke, bench (+ per-kernel table) for every single-GPU workload, reference + eager arms, ncu launch list
#   bash tools/gpu_r2_final.sh full     ncu --set full capture of the hot kernels
mkdir -p gpurun_out
T="timeout 900"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
if [ "$1" = "full" ]; then
  VITB_WGRAD_STREAM=0 $T $CMD > gpurun_out/r2f_plain2.log 2>&1 && \
  VITB_WGRAD_STREAM=0 $T ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_bf16|attn_bwd_bf16|ln_bwd|gelu_bwd|ln_fwd|adam|reduce_jobs" -s ${2:-300} -c ${3:-48} -o /tmp/prof_r2 $CMD > gpurun_out/r2f_ncu_full.log 2>&1
  tail -n 3 gpurun_out/r2f_ncu_full.log | cut -c1-300
  ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_full_raw.csv 2> gpurun_out/r2f_ncu_export.log
  ls -la /tmp/prof_r2.ncu-rep gpurun_out/r2f_ncu_full_raw.csv
  [ $(stat -c %s /tmp/prof_r2.ncu-rep) -lt 40000000 ] && cp /tmp/prof_r2.ncu-rep gpurun_out/
  exit 0
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2f_gpu.txt 2>&1
$T python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2f_tests.log 2>&1
$T python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1
$T python bench.py --kernel-table gpurun_out/r2f_ktable_b1024.json > gpurun_out/r2f_bench.log 2>&1
$T python bench.py --batch 128 --kernel-table gpurun_out/r2f_ktable_b128.json > gpurun_out/r2f_bench_b128.log 2>&1
$T python bench.py --workload t17c100 --no-cpu-baseline --kernel-table gpurun_out/r2f_ktable_t17.json > gpurun_out/r2f_bench_t17c100.log 2>&1
$T python bench.py --workload t17c100 --batch 128 --no-cpu-baseline > gpurun_out/r2f_bench_t17c100_b128.log 2>&1
$T python bench.py --workload scaled65 --no-cpu-baseline > gpurun_out/r2f_bench_scaled65.log 2>&1
$T python bench.py --workload scaled17 --no-cpu-baseline > gpurun_out/r2f_bench_scaled17.log 2>&1
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.log 2>&1
$T python bench.py --impl eager --steps 10 > gpurun_out/r2f_eager_b1024.log 2>&1
VITB_WGRAD_STREAM=0 $T $CMD > gpurun_out/r2f_plain.log 2>&1 && \
VITB_WGRAD_STREAM=0 $T ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launch_summary.txt 2>&1
for f in gpurun_out/r2f_tests.log gpurun_out/r2f_smoke.log gpurun_out/r2f_bench.log gpurun_out/r2f_bench_b128.log gpurun_out/r2f_bench_ref.log gpurun_out/r2f_ncu_list.log; do echo "== $f"; tail -n 3 $f | cut -c1-500; done
head -n 14 gpurun_out/r2f_launch_summary.txt
