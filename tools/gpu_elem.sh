#!/bin/bash
# LayerNorm / GELU backward family: tests, bench, then per-launch time and DRAM bytes from ncu.
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "gelu or layernorm" > gpurun_out/t_elem.log 2>&1; tail -n 3 gpurun_out/t_elem.log
$T python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-200
CMD="python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline"
$T ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gelu_bwd|ln_bwd|partials_finalize" -s 130 -c 50 --csv --log-file gpurun_out/elem_metrics.csv $CMD > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/elem_metrics.csv')) if len(r)>10]
h=rows[0]
cur={}
for r in rows[1:]:
    k=(r[h.index('ID')], r[h.index('Kernel Name')][:60], r[h.index('Grid Size')] if 'Grid Size' in h else '')
    cur.setdefault(k,{})[r[h.index('Metric Name')]]=r[h.index('Metric Value')]
for k,v in cur.items():
    print(k[1], k[2], ' '.join(f"{a.split('__')[-1]}={b}" for a,b in v.items()))
P
