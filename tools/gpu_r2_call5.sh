#!/bin/bash
# Round 2, call 5: single-slab pre-activation epilogue (3 ring stages everywhere), PDL at small batch.
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c5_tests.log 2>&1
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --kernel-table gpurun_out/r2c5_ktable_b1024.json > gpurun_out/r2c5_b1024.log 2>&1
$T $B --batch 128 --kernel-table gpurun_out/r2c5_ktable_b128.json > gpurun_out/r2c5_b128.log 2>&1
VITB_PDL=1 $T $B --batch 128 > gpurun_out/r2c5_b128_pdl.log 2>&1
VITB_PDL=1 $T $B > gpurun_out/r2c5_b1024_pdl.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c5_t17.log 2>&1
VITB_PDL=1 $T $B --workload t17c100 > gpurun_out/r2c5_t17_pdl.log 2>&1
$T $B --workload scaled65 > gpurun_out/r2c5_scaled65.log 2>&1
for f in gpurun_out/r2c5_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-330; done
