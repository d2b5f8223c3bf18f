#!/bin/bash
# Round-end evidence: plain run, ncu launch list of the same command, one `--set full` capture of the dominant kernel.
# Usage: gpurun -- bash scripts/gpu_profile_final.sh <tag>
TAG=${1:-final}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 1 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 120 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ru_ -s 24 -c 6 -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
tail -n 2 $OUT/plain.log | cut -c1-300
