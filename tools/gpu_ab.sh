#!/bin/bash
# A/B visit: parity tests, then one short bench line per environment combination.
# Usage: bash tools/gpu_ab.sh <tag> "<pytest -k expr | all | skip>" "A=1,B=2" "A=0" ...
set -u
TAG=${1:-ab}; KEXPR=${2:-all}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ "$KEXPR" != "skip" ]; then
  if [ "$KEXPR" != "all" ]; then
    timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/test.log 2>&1
  else
    timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test.log 2>&1
  fi
  echo "pytest exit $?" >> $OUT/test.log
  tail -15 $OUT/test.log
fi
i=0
for combo in "$@"; do
  i=$((i+1))
  envs=$(echo "$combo" | tr ',' ' ')
  [ "$combo" = "-" ] && envs=""
  env $envs timeout 600 python bench.py --steps ${AB_STEPS:-60} --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$i.json"))
    print("[$combo] fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "seq_ms", round(d["sequential_ms_per_step"],3),
          "clk", d["clocks"].get("sm_mhz"), " ".join(f'{s["stage"]}={s["ms"]:.3f}' for s in d["roofline_stages"]))
except Exception as e:
    print("[$combo] no bench line:", e)
    import subprocess; print(subprocess.run(["tail","-5","$OUT/bench_$i.err"],capture_output=True,text=True).stdout)
PY
done
