#!/bin/bash
# Tuning visit: parity tests, then one short bench line per value of an env-selected kernel variant.
# Usage: bash tools/gpu_variants.sh <tag> <ENV_VAR> "<values>" [pytest -k expr | skip]
set -u
TAG=${1:-var}; VAR=${2:-PG_COMP_VARIANT}; VALS=${3:-"0"}; KEXPR=${4:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$KEXPR" != "skip" ]; then
  if [ -n "$KEXPR" ]; then
    timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/test.log 2>&1
  else
    timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test.log 2>&1
  fi
  echo "pytest exit $?" >> $OUT/test.log
  tail -15 $OUT/test.log
fi
for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$v.json"))
    print("$VAR=$v fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "seq_ms", round(d["sequential_ms_per_step"],3),
          " ".join(f'{s["stage"]}={s["ms"]:.3f}' for s in d["roofline_stages"]))
except Exception as e:
    print("$VAR=$v no bench line:", e)
PY
done
