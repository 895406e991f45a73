#!/bin/bash
# Quick GPU visit: parity tests + one bench line.  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "${2:-}" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/test.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test.log 2>&1
fi
echo "pytest exit $?" >> $OUT/test.log
tail -15 $OUT/test.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -5 $OUT/bench.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json"))
    print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"].get("sm_mhz"))
    for s in d["roofline_stages"]: print(" ", s["stage"], round(s["ms"],4), round(s.get("frac") or 0,3))
    print(" stats", d["compositing_stats"])
except Exception as e:
    print("no bench line:", e)
PY
