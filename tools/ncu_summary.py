#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) into the few lines kept under profiles/.
usage: ncu_summary.py REPORT.ncu-rep [--stalls]   (reads it with `ncu -i ... --page raw --csv`)"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
        "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    stalls = "--stalls" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## " + r[hdr.index("Kernel Name")])
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                print(f"{h:<75} {v} {u}")
            elif stalls and h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(v) >= 0.05:
                        print(f"{h:<75} {v}")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main()
