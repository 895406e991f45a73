#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, ncu --set full of every pegasus kernel of one frame.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ "${2:-}" != "skip_tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test.log 2>&1
  echo "pytest exit $?" >> $OUT/test.log
  tail -3 $OUT/test.log
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -c 400 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "reference exit $?"; tail -c 300 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > $OUT/launches_run.log 2>&1
# frames before the timed ones: 3 calibration + 3 stats + 3 slot sizing + 1 warm-up = 10 frames x 13 pegasus kernels
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'composite|emit|onesweep|preprocess|hist_kernel|scan_rows|tile_scan|tile_order' -s 130 -c 13 -o $OUT/prof \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > $OUT/prof_run.log 2>&1
python tools/ncu_summary.py $OUT/prof.ncu-rep $OUT/ncu_full_summary.json > /dev/null 2>&1
ls -la $OUT
