#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --batch 128"
VITB_WGRAD_STREAM=0 $T $CMD > gpurun_out/r2c17_plain.log 2>&1 && \
VITB_WGRAD_STREAM=0 $T ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2c17_launches_b128.csv $CMD > gpurun_out/r2c17_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2c17_launches_b128.csv > gpurun_out/r2c17_launch_summary_b128.txt 2>&1
head -n 40 gpurun_out/r2c17_launch_summary_b128.txt
