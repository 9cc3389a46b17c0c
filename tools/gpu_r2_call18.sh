#!/bin/bash
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c18_tests.log 2>&1; tail -n 2 gpurun_out/r2c18_tests.log
B="python bench.py --no-cpu-baseline --steps 30"
for w in "--batch 128" "--workload t17c100 --batch 128" "--batch 256" "--workload t17c100"; do
  tag=$(echo "$w" | tr -d ' -' )
  $T $B $w > gpurun_out/r2c18_${tag}_deep.log 2>&1
  VITB_GEMM_SMALL_DEEP=0 $T $B $w > gpurun_out/r2c18_${tag}_dual.log 2>&1
done
for f in gpurun_out/r2c18_*_*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
timeout 300 python tools/cublas_shapes.py 8320 2>&1 | grep -v '^{' | head -12
