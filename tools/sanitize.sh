#!/bin/bash
# compute-sanitizer over the hot path: memcheck, synccheck and racecheck of __graft_entry__.smoke() (one small
# composed frame in both numerics modes, checked against the oracle) and memcheck of the render-call parity tests.
# Usage (under gpurun): bash tools/sanitize.sh [outdir]      -> <outdir>/sanitize_<tool>.log + sanitize_summary.txt
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p $OUT
SAN=${SANITIZER:-compute-sanitizer}
: > $OUT/sanitize_summary.txt
for tool in memcheck synccheck racecheck; do
  timeout 900 $SAN --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitize_$tool.log 2>&1
  echo "$tool smoke(): exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_$tool.log | tail -1)" >> $OUT/sanitize_summary.txt
done
timeout 1500 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_render_call.py tests/test_gpu_generate.py tests/test_gpu_png.py -m gpu -x -q \
  > $OUT/sanitize_memcheck_tests.log 2>&1
echo "memcheck tests/test_gpu_render_call.py tests/test_gpu_generate.py tests/test_gpu_png.py: exit $? | $(grep -E 'ERROR SUMMARY' $OUT/sanitize_memcheck_tests.log | tail -1) | $(grep -E 'passed|failed' $OUT/sanitize_memcheck_tests.log | tail -1)" >> $OUT/sanitize_summary.txt
timeout 900 $SAN --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_png.py -m gpu -x -q -k "bit_for_bit or batch or strided" \
  > $OUT/sanitize_racecheck_png.log 2>&1
echo "racecheck tests/test_gpu_png.py (small images): exit $? | $(grep -E 'RACECHECK SUMMARY' $OUT/sanitize_racecheck_png.log | tail -1) | $(grep -E 'passed|failed' $OUT/sanitize_racecheck_png.log | tail -1)" >> $OUT/sanitize_summary.txt
cat $OUT/sanitize_summary.txt
