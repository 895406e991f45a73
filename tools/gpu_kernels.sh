#!/bin/bash
# Per-kernel durations of ONE frame (ncu launch list, cold-cache and serialised: compare shares, not absolutes).
# Usage: bash tools/gpu_kernels.sh <tag> [extra env...]
set -u
TAG=${1:-kern}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"preprocess|compact|onesweep|count_kernel|pair_scan|emit|tile_scan|tile_order|scan_rows|composite|pose|pack" \
  -s 300 -c 40 --csv --log-file $OUT/launches.csv python tools/stage_times.py --masks 1 --frames 8 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seen=0
for r in rows[1:]:
    name=r[ki].split("(")[0][:48]
    if "preprocess" in name: seen+=1
    if seen==2: print(f"{name:50s} {float(r[vi])/1000:8.1f} us")
PY
