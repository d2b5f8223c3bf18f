#!/bin/bash
# Quick GPU loop: selected tests (bounded by timeout) + short bench.  Usage: gpurun -- bash scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -q -m gpu -k "$KEXPR" > $OUT/pytest.log 2>&1
else
  timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest.log 2>&1
fi
echo "pytest rc=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|error|Error|mismatch|timeout|max-abs|assert" $OUT/pytest.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke.log
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"])
    r=d["roofline"]; print("roofline",r["kernel"],r["achieved"],r["frac"],r["share_of_step"])
    for k,v in r["classes"].items(): print("  ",k,round(v["ms_per_step"],3),"ms",v["launches_per_step"],"launches",round(v["tflops"],1),"TF",round(v["gbs"],1),"GB/s")
    print("cpu",d["cpu_baseline"]); print("lat",d["latency"]); print("clocks",d["clocks"])
except Exception as e: print("bench parse failed",e)
PY
tail -5 $OUT/bench.err
