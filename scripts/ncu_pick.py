#!/usr/bin/env python
"""Prints selected metrics of an ncu raw-page CSV export (ncu -i X.ncu-rep --page raw --csv > X.csv), one kernel per column."""
import csv
import sys

PICK = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size", "launch__occupancy_limit", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_uniform",
        "gpc__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__cycles_active.avg"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    extra = sys.argv[2:]
    h, u = rows[0], rows[1]
    for i, name in enumerate(h):
        if any(name.startswith(p) or p in name for p in PICK + extra):
            vals = [r[i] for r in rows[2:]]
            if all(v in ("", "0", "n/a") for v in vals) and name != "Kernel Name":
                continue
            print(f"{name[:95]:95s} {u[i]:10s} " + " | ".join(v[:60] for v in vals))


if __name__ == "__main__":
    main()
