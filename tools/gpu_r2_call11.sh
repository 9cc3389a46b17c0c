#!/bin/bash
# Round 2, call 11: which round-2 default costs the scaled ViT (12L/768/3072) its throughput?
mkdir -p gpurun_out
T="timeout 900"
B="python bench.py --no-cpu-baseline --steps 20 --workload scaled17"
$T $B > gpurun_out/r2c11_s17_default.log 2>&1
VITB_WGRAD_STREAM=0 $T $B > gpurun_out/r2c11_s17_nostream.log 2>&1
VITB_PDL=0 $T $B > gpurun_out/r2c11_s17_nopdl.log 2>&1
VITB_DEFER=0 $T $B > gpurun_out/r2c11_s17_nodefer.log 2>&1
VITB_WGRAD_STREAM=1 $T $B > gpurun_out/r2c11_s17_stream1.log 2>&1
B="python bench.py --no-cpu-baseline --steps 20 --workload scaled65"
$T $B --kernel-table gpurun_out/r2c11_ktable_s65.json > gpurun_out/r2c11_s65_default.log 2>&1
VITB_WGRAD_STREAM=0 $T $B > gpurun_out/r2c11_s65_nostream.log 2>&1
for f in gpurun_out/r2c11_s*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c11_ktable_s65.json | head -16
