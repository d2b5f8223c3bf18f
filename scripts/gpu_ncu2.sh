#!/bin/bash
# One gpurun call: plain run, then one `ncu --set full` capture of the kernels matching a regex.
# Usage: gpurun -- bash scripts/gpu_ncu2.sh <tag> <kernel-regex> <skip> <count> [extra bench args]
TAG=$1; KRE=$2; FSKIP=$3; FCNT=$4; shift 4
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 1 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 $@"
$CMD > $OUT/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $FSKIP -c $FCNT -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 $OUT/plain.log $OUT/ncu_full.log
ls -la $OUT
