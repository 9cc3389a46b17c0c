#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c16_tests.log 2>&1; tail -n 2 gpurun_out/r2c16_tests.log
B="python bench.py --no-cpu-baseline --steps 30"
for w in "" "--batch 128" "--workload t17c100" "--workload t17c100 --batch 128"; do
  tag=$(echo "$w" | tr -d ' -' ); tag=${tag:-b1024}
  $T $B $w > gpurun_out/r2c16_${tag}_optlayer.log 2>&1
  VITB_OPT_PER_LAYER=0 $T $B $w > gpurun_out/r2c16_${tag}_optend.log 2>&1
done
for f in gpurun_out/r2c16_*_*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
