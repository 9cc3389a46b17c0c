#!/bin/bash
# ncu launch list (device time of every kernel of one eager step; cold-cache, serialised: compare SHARES)
mkdir -p gpurun_out
T="timeout 900"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$T $CMD > gpurun_out/plain.log 2>&1 && \
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -n 2 gpurun_out/ncu_list.log | cut -c1-300
wc -l gpurun_out/launches.csv
