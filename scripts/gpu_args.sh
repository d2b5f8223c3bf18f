#!/bin/bash
# Bench with extra bench.py arguments.  Usage: gpurun -- bash scripts/gpu_args.sh <tag> "<args A>" ["<args B>" ...]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for A in "$@"; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --latency-reps 3 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 $A > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$i.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("[$A] value %.0f ms/step %.3f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), {k:round(x["ms_per_step"],2) for k,x in r["classes"].items()})
except Exception as e: print("[$A] failed", e)
PY
  i=$((i+1))
done
