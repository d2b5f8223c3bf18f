#!/bin/bash
# One gpurun call: plain run, then ncu launch list + one full capture of a kernel.
# Usage: gpurun -- bash scripts/gpu_ncu.sh <tag> <kernel-regex> [skip] [count] [extra bench args]
TAG=${1:-ncu}; KRE=${2:-k_gemm}; SKIP=${3:-2000}; CNT=${4:-700}; shift 4
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --streams 256 --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 3 $@"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s ${FSKIP:-24} -c ${FCNT:-3} -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 $OUT/ncu_list.log $OUT/ncu_full.log
ls -la $OUT
