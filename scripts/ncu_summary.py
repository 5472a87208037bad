#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py gpurun_out/x.ncu-rep [pattern ...]"""
import csv
import io
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64", "pipe_fp64",
           "sm__pipe_fma", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ", "sm__throughput.avg.pct",
           "gpu__dram_throughput.avg.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op", "l1tex__t_requests_pipe_lsu_mem_global_op",
           "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "smsp__thread_inst_executed_per_inst_executed",
           "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "sm__cycles_elapsed.max", "op_dfma", "op_dadd", "op_dmul",
           "smsp__cycles_active.avg", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "dram__throughput", "smsp__warps_eligible", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed.sum", "shared"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = dict(zip(hdr, r)).get("Kernel Name", "?")
        print(f"== {name[:100]}")
        for h, u, v in zip(hdr, units, r):
            if any(p in h for p in pats):
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main()
