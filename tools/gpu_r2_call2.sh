#!/bin/bash
# Round 2, call 2: full GPU suite on the deferred-reduction / side-stream build, then A/B benches.
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c2_tests.log 2>&1
B="python bench.py --no-cpu-baseline --steps 30"
VITB_DEFER=0 $T $B > gpurun_out/r2c2_b1024_nodefer.log 2>&1
VITB_WGRAD_STREAM=0 $T $B > gpurun_out/r2c2_b1024_defer.log 2>&1
VITB_WGRAD_STREAM=1 $T $B --kernel-table gpurun_out/r2c2_ktable_b1024.json > gpurun_out/r2c2_b1024_side.log 2>&1
VITB_WGRAD_STREAM=2 $T $B > gpurun_out/r2c2_b1024_side_prio.log 2>&1
VITB_DEFER=0 $T $B --batch 128 > gpurun_out/r2c2_b128_nodefer.log 2>&1
VITB_WGRAD_STREAM=0 $T $B --batch 128 > gpurun_out/r2c2_b128_defer.log 2>&1
VITB_WGRAD_STREAM=1 $T $B --batch 128 > gpurun_out/r2c2_b128_side.log 2>&1
VITB_WGRAD_STREAM=2 $T $B --batch 128 > gpurun_out/r2c2_b128_side_prio.log 2>&1
VITB_WGRAD_STREAM=1 $T $B --workload t17c100 > gpurun_out/r2c2_t17_side.log 2>&1
for f in gpurun_out/r2c2_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-400; done
