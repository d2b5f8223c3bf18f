#!/usr/bin/env python
"""profiles/traffic.json from a one-tick `ncu --set full` capture (scripts/gpu_profile_r02.sh): DRAM bytes per window
and launch for every kernel class of bench.py's roofline.  Usage: python scripts/traffic_from_ncu.py <rep> <windows> <src note>"""
import collections
import csv
import io
import json
import subprocess
import sys

CLASS_OF = [("k_deinterleave", "deinterleave"), ("k_codes_head", "from_codes"), ("k_from_codes", "from_codes"),
            ("k_dw_", "dwconv_snake"), ("k_gemm_tc", "gemm_1x1"), ("k_gemm_ws", "gemm_1x1"), ("k_convt", "gemm_convt"),
            ("k_ru_w", "ru_fused_tmem"), ("k_ru_", "ru_fused"), ("k_blk_tail", "block_fused_tail"), ("k_tail", "tail_pack")]


def main():
    rep, windows, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]

    def val(r, k, scale=None):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]]
        if scale == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        if scale == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return v

    detail, agg = [], collections.defaultdict(lambda: [0, 0.0])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        cls = next((c for pat, c in CLASS_OF if pat in name), "other")
        rd, wr = val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB")
        detail.append({"kernel": name, "class": cls, "dram_read_MB": round(rd, 1), "dram_write_MB": round(wr, 1),
                       "us": round(val(r, "gpu__time_duration.sum", "us"), 1),
                       "issue_active_pct": round(val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"), 1),
                       "tensor_active_pct": round(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                       "dram_pct": round(val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                       "regs": int(val(r, "launch__registers_per_thread"))})
        agg[cls][0] += 1
        agg[cls][1] += (rd + wr) * 1e6
    out = {"source": note,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of every launch of one tick, averaged per class over its launches "
                   f"and divided by the {windows} windows of the tick",
           "dram_bytes_per_window_per_launch": {c: v[1] / v[0] / windows for c, v in agg.items()},
           "launches_per_tick": {c: v[0] for c, v in agg.items()},
           "dram_bytes_per_tick": sum(v[1] for v in agg.values()),
           "detail": detail}
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("dram_bytes_per_window_per_launch", "launches_per_tick", "dram_bytes_per_tick")}, indent=1))


if __name__ == "__main__":
    main()
