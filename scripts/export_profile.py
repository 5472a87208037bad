#!/usr/bin/env python
"""Export a tracked text summary of an .ncu-rep into profiles/: python scripts/export_profile.py rep out.txt [title]"""
import subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum ", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fp64.sum ", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread ",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit", "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum ",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum ", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum ",
        "sm__cycles_active.avg ", "sm__cycles_elapsed.max "]
s1 = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep] + [k.strip() for k in KEYS], capture_output=True, text=True).stdout
s1 = "\n".join(l for l in s1.splitlines() if "dshared" not in l and ".peak_sustained" not in l)
s2 = subprocess.run([sys.executable, "scripts/ncu_groups.py", rep, "8"], capture_output=True, text=True).stdout
s3 = subprocess.run([sys.executable, "scripts/ncu_source_top.py", rep, "25"], capture_output=True, text=True).stdout
open(out, "w").write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on (one launch), read with ncu -i --page raw/source --csv\n\n"
                     f"## raw metrics\n{s1}\n\n## SASS grouped by execution count (loop nest): instructions, warp-level instruction total, stall samples\n{s2}\n"
                     f"## top stalled SASS instructions\n{s3}\n")
print(out)
