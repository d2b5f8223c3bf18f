"""Per-kernel counts of the Blackwell-native SASS opcodes in libsnacb.so (cuobjdump -sass): tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP/UBLKPF, packed fp32 -> FFMA2/FMUL2/FADD2, legacy HMMA must be absent.
Usage: python scripts/sass_opcodes.py [lib] > profiles/r02_sass_opcodes.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "project_morpheus_b200/libsnacb.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTCBAR", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "LDGSTS",
         "HMMA", "HGMMA"]
fn, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("(anonymous namespace)::", "").replace("snacb::", "").replace("void ", "")
        fn = re.sub(r"\(CUtensorMap.*|\((?:snacb|int|float|unsigned|const|long|__half).*", "", fn).replace("(int)", "").replace("(bool)", "")
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_all"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[fn][w] += 1
                total[w] += 1
print("# SASS opcode evidence (round 2)\n")
print(f"`cuobjdump -sass {lib}` aggregated by `scripts/sass_opcodes.py`; static instruction counts per kernel.\n")
print("Totals: " + ", ".join(f"{w} {total[w]}" for w in WATCH) + "\n")
cols = [w for w in WATCH if total[w] or w in ("HMMA", "HGMMA")]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for fn, c in counts.items():
    if not any(c[w] for w in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKPF", "FFMA2", "MUFU", "LDGSTS")):
        continue
    print(f"| `{fn}` | {c['_all']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
