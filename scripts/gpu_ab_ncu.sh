#!/bin/bash
# Per-kernel times (ncu launch list of one tick) for several values of an env var.  Usage: gpurun -- bash scripts/gpu_ab_ncu.sh <tag> VAR "v1 v2" [kernel regex]
TAG=$1; VAR=$2; VALS=$3; RE=${4:-k_}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 1 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 --sustained-s 0 --cfg4-streams 0 --pull-streams="
export SNACB_PROFILE_STEP=1
for v in $VALS; do
  env $VAR=$v ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:$RE --csv --log-file $OUT/list_$v.csv $CMD > $OUT/run_$v.log 2>&1
  echo "== $VAR=$v"; python - <<PY
import csv,io
t=open("$OUT/list_$v.csv").read(); t=t[t.index('"ID"'):]
for r in csv.DictReader(io.StringIO(t)):
    if r["Metric Name"]=="gpu__time_duration.sum": print("  ", r["Kernel Name"].split("(")[0][-34:], r["Metric Value"], r["Metric Unit"])
PY
done
