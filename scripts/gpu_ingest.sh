#!/bin/bash
# Token-ingress leg of bench.py only (plus the GPU test that drives the native scheduler).  Usage: gpurun -- bash scripts/gpu_ingest.sh <tag>
TAG=${1:-ingest}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests -q -m gpu -x -k "scheduler or ingest" 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-reps 3 --long-read-batch 0 --ragged-streams 0 > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(json.dumps(d["latency"]["n2_token_ingress"], indent=1))
PY
tail -3 $OUT/bench.err
