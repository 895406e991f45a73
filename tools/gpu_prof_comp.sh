#!/bin/bash
# ncu --set full of ONE compositing launch of a timed frame per PG_COMP_VARIANT value.
# Usage: bash tools/gpu_prof_comp.sh <tag> "<variants>"
set -u
TAG=${1:-profc}; VARS=${2:-"4"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARS; do
  # launches before the timed frames of `--steps 2 --warmup 1`: 3 calibration + 3 stats + 3 slot sizing + 1 warm-up
  PG_COMP_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:composite -s 10 -c 1 \
    -o $OUT/comp_v$v python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > $OUT/prof_v$v.log 2>&1
  python tools/ncu_summary.py $OUT/comp_v$v.ncu-rep $OUT/comp_v${v}_summary.json > /dev/null 2>&1
done
ls -la $OUT
