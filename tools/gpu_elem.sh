#!/bin/bash
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "gelu or layernorm" > gpurun_out/t_elem.log 2>&1; tail -n 3 gpurun_out/t_elem.log
$T python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-200
CMD="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline"
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gelu_bwd|rows_colsum|ln_fwd|ln_bwd" -s 60 -c 12 --csv --log-file gpurun_out/elem_metrics.csv $CMD > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/elem_metrics.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
P
