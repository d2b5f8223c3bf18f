#!/bin/bash
# ncu launch list only.  Usage: gpurun -- bash scripts/gpu_list.sh <tag> [skip] [count]
TAG=${1:-list}; SKIP=${2:-200}; CNT=${3:-100}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --streams 256 --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 3"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "list rc=$?"
