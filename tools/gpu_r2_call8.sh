#!/bin/bash
# Round 2, call 8: early PDL trigger for single-wave grids (default build) vs none (build/libvitb200_noearly.so)
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c8_tests.log 2>&1
B="python bench.py --no-cpu-baseline --steps 30"
NE=$PWD/vit-cifar_b200/build/libvitb200_noearly.so
for w in "" "--batch 128" "--workload t17c100" "--workload t17c100 --batch 128"; do
  tag=$(echo "$w" | tr -d ' -' ); tag=${tag:-b1024}
  $T $B $w > gpurun_out/r2c8_${tag}_early.log 2>&1
  VITB_LIB_PATH=$NE $T $B $w > gpurun_out/r2c8_${tag}_noearly.log 2>&1
done
for f in gpurun_out/r2c8_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-260; done
