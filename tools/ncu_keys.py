#!/usr/bin/env python
"""Prints the headline metrics of one kernel from an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
kfilter = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_uniform.sum",
        "smsp__inst_executed_pipe_cbu.sum", "smsp__inst_executed_pipe_adu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_global_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    if kfilter and kfilter not in name:
        continue
    print("== kernel", name[:100])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print("  %-70s %s %s" % (h, r[i], rows[1][i]))
        elif "average_warp_latency_issue_stalled" in h or ("issue_stalled" in h and h.endswith("per_warp_active.pct")):
            try:
                if float(r[i]) > 1.0:
                    print("  %-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall_"), r[i]))
            except ValueError:
                pass
