#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "attention" > gpurun_out/t_attn.log 2>&1; tail -n 3 gpurun_out/t_attn.log
$T python bench.py --kernel-table gpurun_out/kernels_b1024.json --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-200
python - <<'P'
import json
d=json.load(open('gpurun_out/kernels_b1024.json'))
for k in d['kernels']:
    if k['op'].startswith('attn'): print(k['op'], round(k['ms_per_call']*1e3,1), 'us')
P
CMD="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline"
$T ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio --clock-control none -k regex:"attn_" -s 40 -c 3 --csv --log-file gpurun_out/attn_metrics.csv $CMD > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/attn_metrics.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:30], r[h.index('Metric Name')][-60:], r[h.index('Metric Value')])
P
