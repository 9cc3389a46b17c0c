#!/bin/bash
# Round 2, call 6: LayerNorm forward with 4 rows in flight per warp; defaults (PDL, priority stream) in the test suite.
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/r2c6_tests.log 2>&1
B="python bench.py --no-cpu-baseline --steps 30"
$T $B --kernel-table gpurun_out/r2c6_ktable_b1024.json > gpurun_out/r2c6_b1024.log 2>&1
$T $B --batch 128 --kernel-table gpurun_out/r2c6_ktable_b128.json > gpurun_out/r2c6_b128.log 2>&1
$T $B --workload t17c100 > gpurun_out/r2c6_t17.log 2>&1
for f in gpurun_out/r2c6_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-330; done
python tools/ktable.py gpurun_out/r2c6_ktable_b1024.json | head -22
