#!/usr/bin/env python
"""Summarise `ncu --set full` reports (.ncu-rep) into profiles/<name>.json.

usage: tools/ncu_summary.py OUT.json label=report.ncu-rep [label=report.ncu-rep ...]

Keeps what DESIGN.md / bench.py quote: duration, DRAM bytes, pipe utilisations
(sm__inst_executed_pipe_*), issue rate, stall reasons, registers, occupancy, and the
dynamic instruction count."""
import csv
import io
import json
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
)


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, unit, val = rows[0], rows[1], rows[2]
    d = {}
    for k, u, v in zip(head, unit, val):
        if k in KEEP or k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") \
                or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
            d[k] = ("%s %s" % (v, u)).strip()
        if k == "Kernel Name":
            d["kernel"] = v
    return d


if __name__ == "__main__":
    res = {}
    for arg in sys.argv[2:]:
        label, path = arg.split("=", 1)
        res[label] = summarise(path)
    json.dump(res, open(sys.argv[1], "w"), indent=1, sort_keys=True)
    print("wrote", sys.argv[1], list(res))
