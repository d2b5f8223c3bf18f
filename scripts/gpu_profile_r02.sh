#!/bin/bash
# Round-2 evidence: plain run, then the ncu launch list and one `--set full` capture of exactly ONE tick of the same
# command (bench.py brackets its last device-timed step with cudaProfilerStart/Stop when SNACB_PROFILE_STEP=1).
# Usage: gpurun -- bash scripts/gpu_profile_r02.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-reps 1 --long-read-batch 0 --ragged-streams 0 --ingest-streams 0 --sustained-s 0 --cfg4-streams 0 --pull-streams="
export SNACB_PROFILE_STEP=1
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"
tail -n 1 $OUT/plain.log | cut -c1-400
grep -c gpu__time_duration $OUT/launches.csv
