#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --kernel-table gpurun_out/r2c36_ktable_b1024.json > gpurun_out/r2c36_bench.log 2>&1
timeout 200 python bench.py --batch 128 --no-cpu-baseline > gpurun_out/r2c36_bench_b128.log 2>&1
for f in gpurun_out/r2c36_bench*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-220; done
