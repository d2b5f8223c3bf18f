#!/bin/bash
# Sweep bench.py over chunk sizes.  Usage: gpurun -- bash scripts/gpu_sweep.sh <tag> "<chunks>"
TAG=${1:-sweep}; CHUNKS=${2:-"32 128 512 1024"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in $CHUNKS; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --latency-reps 3 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 --chunk $c > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_c$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("chunk $c value %.0f ms/step %.2f e2e %.0f launches/step %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]/5), {k:round(v["ms_per_step"],2) for k,v in r["classes"].items()})
except Exception as e: print("chunk $c failed", e)
PY
done
