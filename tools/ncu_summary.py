#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into the handful of counters DESIGN.md / profiles/ quote.

usage: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [title] > profiles/<name>.md
Reads the report with `ncu -i ... --page raw --csv` (works on a box without a GPU).
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "blocks/SM allowed by registers"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of max warps)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used (%)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe busy (%) [IMAD lives here]"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe instructions (% of peak)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe instructions (% of peak)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe instructions (% of peak)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate (%)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local-memory store instructions (spills)"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local-memory load instructions (spills)"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("Source: `%s` (`ncu --set full --clock-control none --import-source on`), %d captured launch(es).\n" % (rep, len(data)))
    for d in data:
        print("## `%s`  grid %s block %s\n" % (d[col["Kernel Name"]], d[col["Grid Size"]], d[col["Block Size"]]))
        print("| counter | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in col:
                print("| %s (`%s`) | %s | %s |" % (label, key, d[col[key]], units[col[key]]))
        stalls = []
        for h, i in col.items():
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[i]), h[len(STALLS):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("\nWarp stall reasons (warps stalled per issued instruction, top 8):\n")
        print("| reason | warps/issue |\n|---|---:|")
        for v, name in stalls[:8]:
            print("| %s | %.3f |" % (name, v))
        print()


if __name__ == "__main__":
    main()
