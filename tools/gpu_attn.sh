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
