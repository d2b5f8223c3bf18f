"""Aggregate the executed SASS instructions of one kernel from `ncu --page source --csv` by opcode.
Usage: ncu -i rep --page source --csv --kernel-id :::N > src.csv; python scripts/opcode_hist.py src.csv <elements>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
elems = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = None
ops, samples, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if "Instructions Executed" in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) // 2:
        continue
    toks = r[hdr["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    key = parts[0] + ("." + parts[1] if parts[0] in ("MUFU", "LDG", "STG", "LDS", "STS", "F2F", "F2FP", "UTCHMMA") and len(parts) > 1 else "")
    n = int(r[hdr["Instructions Executed"]])
    ops[key] += n
    tot += n
    samples[key] += int(r[hdr["# Samples"]])
print("total warp instructions", tot)
for k, v in ops.most_common(45):
    extra = f"  thread-inst/elem {v * 32 / elems:5.2f}" if elems else ""
    print(f"{k:16s} {v:12d} {v / tot:6.3f}{extra}  samples {samples[k]}")
