#!/usr/bin/env python
"""Print the handful of raw ncu metrics that decide where the statistics kernel is bound.
usage: ncu_keys.py raw.csv [raw2.csv ...]   (from `ncu -i x.ncu-rep --page raw --csv`)"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__cycles_active.avg"]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    d = dict(zip(rows[0], rows[-1]))
    u = dict(zip(rows[0], rows[1]))
    print(f, d.get("Kernel Name", "")[:70])
    for k in KEYS:
        print(f"   {k} = {d.get(k)} {u.get(k, '')}")
    for k in rows[0]:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            x = float(d[k]) if d[k] else 0
            if x > 0.1:
                print("      stall", k.split("issue_stalled_")[1].split("_per")[0], round(x, 2))
