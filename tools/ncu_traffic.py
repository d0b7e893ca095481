#!/usr/bin/env python
"""Condense an `ncu --set full` report into profiles/: per captured launch the duration, DRAM bytes,
throughput percentages and occupancy (text table), and profiles/traffic.json = DRAM bytes per launch per
kernel (mean over the captured launches), which bench.py reports as roofline.traffic.

usage: tools/ncu_traffic.py gpurun_out/<tag>_prof.ncu-rep <tag>      (runs `ncu -i ... --page raw --csv`)"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("capdec::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(rep, tag):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s (ncu --set full --clock-control none; cold caches, serialised launches)" % os.path.basename(rep),
             "%-44s %-14s %9s %11s %11s %7s %7s %7s %7s %5s %7s" % ("kernel", "grid", "us", "dram_rd_MB", "dram_wr_MB",
                                                                 "dram%", "sm%", "tensor%", "warps%", "regs", "L2hit%")]
    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = short(r[col["Kernel Name"]])
        def g(m):
            return r[col[m]] if m in col else "nan"
        us = float(g(WANT[0]).replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[col[WANT[0]]], 1)
        rd = to_bytes(g(WANT[1]), units[col[WANT[1]]])
        wr = to_bytes(g(WANT[2]), units[col[WANT[2]]])
        lines.append("%-44s %-14s %9.2f %11.3f %11.3f %7.1f %7.1f %7.1f %7.1f %5s %7.1f" % (
            k[:44], r[col["Grid Size"]].replace(" ", ""), us, rd / 1e6, wr / 1e6, float(g(WANT[3])), float(g(WANT[4])),
            float(g(WANT[5])), float(g(WANT[6])), g(WANT[7]), float(g(WANT[8]))))
        key = re.sub(r"<.*", "", k)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += rd + wr
        a[2] += us
    with open(os.path.join(ROOT, "profiles", "%s_full_summary.txt" % tag), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        traffic = json.load(open(path))
    except Exception:
        traffic = {}
    for k, (n, b, us) in agg.items():
        traffic[k] = {"dram_bytes_per_launch": b / n, "us_per_launch_under_ncu": us / n, "launches": n,
                      "source": "profiles/%s_full_summary.txt" % tag}
    if "attn_scores_kernel" in traffic and "attn_wsum_kernel" in traffic:
        traffic["attn_scores_kernel+attn_wsum_kernel"] = {
            "dram_bytes_per_launch": traffic["attn_scores_kernel"]["dram_bytes_per_launch"] +
            traffic["attn_wsum_kernel"]["dram_bytes_per_launch"],
            "source": traffic["attn_wsum_kernel"]["source"]}
    json.dump(traffic, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
