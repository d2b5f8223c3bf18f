#!/usr/bin/env python
"""Summarise a gpurun ncu capture into profiles/<name>.md (launch-list shares + key metrics of the
full capture).  Usage: python scripts/summarize_ncu.py gpurun_out/<tag> profiles/<name>.md "<title>" """
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_xu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launch_table(path):
    txt = open(path).read()
    r = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"]
        k = k.split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} | {v[1] / v[0]:.2f} |")
    return "\n".join(out)


def full_table(rep):
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    except Exception as exc:  # noqa: BLE001
        return f"(could not read {rep}: {exc})"
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        rec = dict(zip(hdr, d))
        out.append(f"**{rec.get('Kernel Name', '?')[:100]}** (id {rec.get('ID')}, grid {rec.get('Grid Size')}, block {rec.get('Block Size')})\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k) for k in KEYS):
                out.append(f"| {h} | {d[i]} | {units[i]} |")
        out.append("")
    return "\n".join(out)


def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    parts = [f"# {title}\n", f"Source: `{src}/launches.csv` (ncu --metrics gpu__time_duration.sum --clock-control none) and "
             f"`{src}/prof.ncu-rep` (ncu --set full).  Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n",
             "## Launch list (aggregated by kernel)\n", launch_table(f"{src}/launches.csv"), "\n## Full capture (selected metrics)\n",
             full_table(f"{src}/prof.ncu-rep")]
    open(dst, "w").write("\n".join(parts) + "\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    main()
