#!/bin/bash
# One ncu --set full capture of the kernels matching a regex during a short bench run.
# Usage: gpurun -- bash scripts/gpu_ncu.sh <tag> <kernel regex> [skip] [count]
TAG=${1:-ncu}; RE=${2:-k_blk}; SKIP=${3:-2}; CNT=${4:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --latency-reps 5 --ingest-streams 0 --ragged-streams 0 --long-read-batch 0 ${BENCH_ARGS:-}"
SNACB_DEBUG=1 python bench.py $ARGS > $OUT/plain.json 2> $OUT/plain.err &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/prof python bench.py $ARGS > $OUT/ncu.log 2>&1
echo "rc=$?"; tail -3 $OUT/plain.err; tail -3 $OUT/ncu.log
