#!/bin/bash
# Sweep bench.py over "chunk:lanes" pairs.  Usage: gpurun -- bash scripts/gpu_sweep2.sh <tag> "64:4 128:4 ..."
TAG=${1:-sweep2}; PAIRS=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for pr in $PAIRS; do
  c=${pr%%:*}; l=${pr##*:}
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --latency-reps 3 --long-read-batch 0 --ragged-streams 0 --chunk $c --lanes $l > $OUT/b_${c}_$l.json 2> $OUT/b_${c}_$l.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/b_${c}_$l.json").read().strip().splitlines()[-1])
    print("chunk $c lanes $l value %.0f ms/step %.3f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]))
except Exception as e: print("chunk $c lanes $l failed", e); print(open("$OUT/b_${c}_$l.err").read()[-600:])
PY
done
