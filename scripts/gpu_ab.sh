#!/bin/bash
# A/B bench over an env var.  Usage: gpurun -- bash scripts/gpu_ab.sh <tag> VAR "v1 v2 ..."
TAG=$1; VAR=$2; VALS=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in $VALS; do
  env $VAR=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --latency-reps 3 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$VAR=$v value %.0f ms/step %.3f e2e %.0f sync %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"].get("sync_value",0)), {k:round(x["ms_per_step"],2) for k,x in r["classes"].items()})
except Exception as e: print("$VAR=$v failed", e)
PY
done
