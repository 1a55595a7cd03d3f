#!/usr/bin/env python
"""Compact summary of an .ncu-rep (one column per distinct kernel, first captured launch of each).

    python tools_ncu_summary.py gpurun_out/x.ncu-rep [name-filter]
"""
import csv
import subprocess
import sys

WANT = [
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_rd_GB", "dram__bytes_read.sum"),
    ("dram_wr_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("l1_wavefronts_M", "l1tex__data_pipe_lsu_wavefronts.sum"),
    ("l2_bytes_GB", "lts__t_bytes.sum"),
]


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("femb::", "")
        if flt and flt not in name:
            continue
        seen.setdefault(name, r)
    for name, r in seen.items():
        print("==", name)
        for label, key in WANT:
            if key in H:
                i = H.index(key)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    if label.endswith("_M"):
                        v = f"{f / 1e6:.1f}"
                    elif U[i] == "byte":
                        v = f"{f / 1e9:.3f} (GB)"
                    else:
                        v = f"{f:.3f}"
                except ValueError:
                    pass
                print(f"   {label:22s} {v} {U[i]}")


if __name__ == "__main__":
    main()
