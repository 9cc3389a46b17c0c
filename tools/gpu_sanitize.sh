#!/bin/bash
# One compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck
# Runs the kernel-level GPU tests (GEMM / attention / LayerNorm / fused backward / reductions) at the small and medium shapes and
# one small engine step under the tool; the plain run comes first (same command line, no tool) as the recipe requires.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='(gemm or attention or layernorm or fused or gelu or colsum or ls_ce or adam or patch or dropout) and not 66560 and not 33280 and not 20000 and not 17408 and not 8320 and not benchmark'
CMD="python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -q -m gpu --timeout 1200 -x -p no:cacheprovider -k"
timeout 900 $CMD "$SEL" > gpurun_out/sanitize_plain.log 2>&1
rc=$?
tail -n 3 gpurun_out/sanitize_plain.log
if [ $rc -ne 0 ]; then echo "plain run failed (rc $rc): not running the sanitizer"; exit 0; fi
timeout 2400 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitize_$TOOL.log $CMD "$SEL" > gpurun_out/sanitize_${TOOL}_pytest.log 2>&1
echo "sanitizer rc $?"
tail -n 4 gpurun_out/sanitize_${TOOL}_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$TOOL.log | tail -n 5
wc -l gpurun_out/sanitize_$TOOL.log
