#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: python profiles/ncu_pick.py raw.csv [pattern ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
           "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
           "launch__waves", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu",
           "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_xu", "sm__inst_executed.sum", "smsp__inst_executed.sum ",
           "sm__pipe_fp64_cycles_active", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active",
           "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed", "lts__t_bytes.sum ",
           "l1tex__t_bytes.sum ", "smsp__warp_issue_stalled", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pats = sys.argv[2:] or DEFAULT
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(p.strip() in h for p in pats) and r[i] not in ("", "0", "n/a"):
                print("  %-90s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
