#!/bin/bash
# One GPU visit for the committed evidence: parity tests, smoke, bench line, reference arm, ncu launch list,
# ncu --set full of every pegasus kernel of one frame (+ the pose / pack kernels).
# Usage (from the repo root, under gpurun): bash tools/gpu_round2.sh <tag> [skip_tests]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ "${2:-}" != "skip_tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/test.log 2>&1
  echo "pytest exit $?" >> $OUT/test.log
  tail -3 $OUT/test.log
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -c 300 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "reference exit $?"; tail -c 300 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > $OUT/launches_run.log 2>&1
# frames before the timed ones with --steps 2 --warmup 1: 3 calibration + 3 stats + 3 slot sizing + 1 warm-up
# = 10 frames x 15 pegasus kernels (preprocess, compact_hist, scan_rows, 4 + 2 onesweep, count, pair_scan, emit,
# tile_scan, tile_order, composite)
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'composite|emit_kernel|count_kernel|pair_scan|onesweep|preprocess|compact_hist|scan_rows|tile_scan|tile_order' -s 150 -c 15 -o $OUT/prof \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > $OUT/prof_run.log 2>&1
python tools/ncu_summary.py $OUT/prof.ncu-rep $OUT/ncu_full_summary.json > $OUT/ncu_full_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pose_kernel|pack_kernel|pack_masks' -s 6 -c 3 -o $OUT/prof_pose \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > $OUT/prof_pose_run.log 2>&1
python tools/ncu_summary.py $OUT/prof_pose.ncu-rep $OUT/ncu_pose_summary.json > $OUT/ncu_pose_summary.txt 2>&1
# pg_png_encode on one frame of the same workload: timing, then ncu --set full of one steady-state launch triple
# (3 kernels per call; the calibration of png_time.py makes 8 calls, 3 warm-up calls follow)
timeout 600 python tools/png_time.py > $OUT/png_time.json 2> $OUT/png_time.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:png_ -s 33 -c 3 -o $OUT/prof_png \
  python tools/png_time.py --n 4 > $OUT/prof_png_run.log 2>&1
python tools/ncu_summary.py $OUT/prof_png.ncu-rep $OUT/ncu_png_summary.json > $OUT/ncu_png_summary.txt 2>&1
ls -la $OUT
