#!/bin/bash
# Like gpu_ncu2.sh with an environment assignment for the profiled command.
# Usage: gpurun -- bash scripts/gpu_ncu3.sh <tag> <VAR=value> <kernel-regex> <skip> <count>
TAG=$1; ENVV=$2; KRE=$3; FSKIP=$4; FCNT=$5
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 1 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0"
env $ENVV $CMD > $OUT/plain.log 2>&1 &&
env $ENVV ncu --set full --clock-control none --import-source on -k regex:$KRE -s $FSKIP -c $FCNT -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
tail -n 3 $OUT/plain.log $OUT/ncu_full.log
