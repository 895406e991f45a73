#!/bin/bash
# Baseline visit: the cross-check tests against baseline/upstream_style.cu + one bench line with gpu_baseline.
set -u
TAG=${1:-base}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_baseline.py -m gpu -q > $OUT/test.log 2>&1
echo "pytest exit $?" >> $OUT/test.log
tail -40 $OUT/test.log
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; tail -5 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1)); print(d["gpu_baseline"])
PY
