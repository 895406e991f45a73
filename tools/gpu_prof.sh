#!/bin/bash
# ncu --set full of the pegasus kernels of ONE timed frame.  Usage: bash tools/gpu_prof.sh <tag> [kernel regex]
set -u
TAG=${1:-prof}
RE=${2:-composite|emit|onesweep|preprocess|hist_kernel|tile_scan}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# --steps 2 --warmup 1: 3 calibration + 3 stats + 1 warm-up frames precede the timed ones; 12 kernels per frame
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s ${3:-84} -c ${4:-12} -o $OUT/prof \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > $OUT/prof_run.log 2>&1
ls -la $OUT
