#!/usr/bin/env python
"""Short summary of one kernel's `ncu --set full` capture: python tools/ncu_brief.py report.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared", "sm__inst_executed_pipe_lsu",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    if sub and sub not in name:
        continue
    print("==", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in WANT or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3):
            print(f"  {h} [{u}] = {v}")
