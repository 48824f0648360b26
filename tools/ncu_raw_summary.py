#!/usr/bin/env python3
"""Aggregate an `ncu --page raw --csv` export (one row per captured launch) per kernel.

usage: python tools/ncu_raw_summary.py gpurun_out/<name>_raw.csv "title" profiles/<name>.json > profiles/<name>.md

The JSON holds, per kernel, the summed duration and DRAM bytes over the captured launches: bench.py reads it
to fill `roofline.traffic` (DRAM bytes per launch of the dominant kernel, from this capture of the same command).
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}  # durations in ms


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pasta::", "")
    return name.strip()


def main():
    path, title = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        try:
            return float(r[col[key]].replace(",", "")) * UNIT.get(units[col[key]], 1.0)
        except (KeyError, ValueError):
            return float("nan")

    agg = collections.OrderedDict()
    for r in data:
        k = short(r[col["Kernel Name"]])
        a = agg.setdefault(k, {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0, "w": collections.defaultdict(float),
                               "regs": r[col["launch__registers_per_thread"]]})
        t = val(r, "gpu__time_duration.sum")
        a["launches"] += 1
        a["ms"] += t
        a["dram_read"] += val(r, "dram__bytes_read.sum")
        a["dram_write"] += val(r, "dram__bytes_write.sum")
        for key in ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
                    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"):
            v = val(r, key)
            if v == v:
                a["w"][key] += v * t  # time-weighted
    total = sum(a["ms"] for a in agg.values())
    print("# %s\n" % title)
    print("Source: `%s` (`ncu --set full --clock-control none`, %d launches, %.2f ms of kernel time; per-launch times under ncu are "
          "cold-cache and serialised -- compare shares).  Percentages are duration-weighted means over a kernel's launches.\n" % (path, len(data), total))
    print("| kernel | launches | regs | ms | share | DRAM read MB | DRAM write MB | DRAM GB/s | L2 hit % | fmaheavy busy % | issue % | occupancy % | alu % | lsu % | stall math_throttle | stall long_sb | stall wait |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    out = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        w = {key: (v / a["ms"] if a["ms"] else 0.0) for key, v in a["w"].items()}
        g = lambda key: w.get(key, float("nan"))
        bw = (a["dram_read"] + a["dram_write"]) / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0.0
        print("| `%s` | %d | %s | %.3f | %.1f %% | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f | %.2f | %.2f |" % (
            k, a["launches"], a["regs"], a["ms"], 100 * a["ms"] / total, a["dram_read"] / 1e6, a["dram_write"] / 1e6, bw,
            g("lts__t_sector_hit_rate.pct"), g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            g("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
            g("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            g("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")))
        out[k] = {"launches": a["launches"], "ms": a["ms"], "dram_read_bytes": a["dram_read"], "dram_write_bytes": a["dram_write"],
                  "fmaheavy_busy_pct": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                  "l2_hit_pct": g("lts__t_sector_hit_rate.pct")}
    if len(sys.argv) > 3:
        json.dump({"source": path, "kernels": out}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
