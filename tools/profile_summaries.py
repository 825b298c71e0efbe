#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/:

  python tools/profile_summaries.py launches gpurun_out/X_launches.csv profiles/rNN_launches.csv
      per-kernel aggregate of an `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/profile_summaries.py full gpurun_out/X.ncu-rep profiles/rNN_kernel_ncu_full.txt "note"
      selected metrics of one `ncu --set full` capture (raw page)
  python tools/profile_summaries.py regions gpurun_out/X.ncu-rep profiles/rNN_regions.txt n_warp_items "note"
      per-loop region table from the source page (tools/ncu_regions.py)
"""
import csv
import io
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active"]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        t = float(r[v].replace(",", "")) / 1e3
        a = agg.setdefault(r[k], [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    ours = sum(a[1] for n, a in agg.items() if "ppm::" in n and "dfma" not in n)
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "max_us", "share_of_ppm_device_time"])
        for n, a in agg.items():
            share = a[1] / ours if ("ppm::" in n and "dfma" not in n) else float("nan")
            w.writerow([n, a[0], round(a[1], 1), round(a[1] / a[0], 1), round(a[2], 1), round(share, 4)])


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def full(rep, dst, note):
    hdr, units, vals = raw_page(rep)
    with open(dst, "w") as f:
        f.write(note + "\n")
        for i, h in enumerate(hdr):
            if h in KEEP or h == "Kernel Name" or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                f.write(f"{h} [{units[i]}] = {vals[i]}\n")


def regions(rep, dst, items, note):
    tmp = dst + ".src.csv"
    with open(tmp, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=f, text=True)
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "ncu_regions.py"), tmp, items], capture_output=True, text=True).stdout
    os.remove(tmp)
    with open(dst, "w") as f:
        f.write("# " + note + "\n# columns: offset of the first instruction, n instructions, FP64 instructions, executions per warp item, share of\n"
                "# executed instructions, share of warp-time samples, share of FP64 instructions, samples/exec ratio, top stall reasons\n" + out)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif cmd == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif cmd == "regions":
        regions(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
