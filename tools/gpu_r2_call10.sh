#!/bin/bash
# Round 2, call 10: attention forward at 5 CTAs per SM (default build) vs 4 (build/libvitb200_fwd4.so)
mkdir -p gpurun_out
T="timeout 900"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 120 -x -k "attention" > gpurun_out/r2c10_tests_attn.log 2>&1
rc=$?; tail -n 3 gpurun_out/r2c10_tests_attn.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "attention tests failed"; exit 0; fi
B="python bench.py --no-cpu-baseline --steps 30"
F4=$PWD/vit-cifar_b200/build/libvitb200_fwd4.so
$T $B --kernel-table gpurun_out/r2c10_ktable_b1024.json > gpurun_out/r2c10_b1024.log 2>&1
VITB_LIB_PATH=$F4 $T $B --kernel-table gpurun_out/r2c10_ktable_b1024_fwd4.json > gpurun_out/r2c10_b1024_fwd4.log 2>&1
$T $B --batch 128 > gpurun_out/r2c10_b128.log 2>&1
VITB_LIB_PATH=$F4 $T $B --batch 128 > gpurun_out/r2c10_b128_fwd4.log 2>&1
for f in gpurun_out/r2c10_b*.log; do echo "== $f"; grep '^{' $f | tail -n 1 | cut -c1-200; done
python tools/ktable.py gpurun_out/r2c10_ktable_b1024.json | grep attn
python tools/ktable.py gpurun_out/r2c10_ktable_b1024_fwd4.json | grep attn
