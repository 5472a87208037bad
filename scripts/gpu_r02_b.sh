#!/bin/bash
# round 2, call B: ncu --set full of the C3 filter at the 8-GPU shard size (N = 8,192: one warp per SM sub-partition)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_small_lw -s 0 -c 1 -o gpurun_out/prof_r02_lw_n8192_all \
  python scripts/strong_probe.py 8192 > gpurun_out/prof_r02_lw_n8192_all.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_small_lw -s 7 -c 1 -o gpurun_out/prof_r02_lw_n8192_ll \
  python scripts/strong_probe.py 8192 > gpurun_out/prof_r02_lw_n8192_ll.log 2>&1
tail -3 gpurun_out/prof_r02_lw_n8192_ll.log
ls -la gpurun_out/*.ncu-rep
