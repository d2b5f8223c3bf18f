"""Region-level view of one kernel from `ncu --page source --csv`: consecutive SASS instructions with the same execution
count are grouped; prints instructions, executions, warp-instructions and stall samples per region.
Usage: ncu -i rep --page source --csv > src.csv; python scripts/ncu_regions.py src.csv [units] [kernel index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out, k = [], -1
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        continue
    if k == which:
        out.append(r)
hdr = {h: i for i, h in enumerate(out[0])}
recs = []
for r in out[1:]:
    if len(r) < 10:
        continue
    recs.append((int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]]), r[hdr["Source"]].strip()))
tot_s = sum(x[1] for x in recs)
tot_i = sum(x[0] for x in recs)
print(f"total warp-instr {tot_i} ({tot_i / units:.1f} per unit), samples {tot_s}")
prev, start, s = None, 0, 0
def flush(i):
    cnt = i - start
    if cnt * prev > 0.004 * tot_i or s > 0.004 * tot_s:
        ops = collections.Counter(x[2].split()[1 if x[2].startswith("@") else 0].split(".")[0] for x in recs[start:i])
        top = ",".join(f"{k}{v}" for k, v in ops.most_common(4))
        print(f"instrs {start:5d}-{i - 1:5d} ({cnt:4d}) exec/unit {prev / units:8.2f} warp-instr/unit {cnt * prev / units:9.1f} "
              f"samples {s:6d} ({100.0 * s / tot_s:4.1f}%)  {top}")
for i, (n, sm, src) in enumerate(recs):
    if prev is None:
        prev, start = n, i
    if abs(n - prev) > prev * 0.02 + 1:
        flush(i)
        prev, start, s = n, i, 0
    s += sm
flush(len(recs))
