#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python scripts/bench_configs.py c4 2>&1 | tee gpurun_out/r02s_c45.json | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/prof_r02_ukf_seg3 python scripts/profile_generic.py ukf > gpurun_out/prof_r02_ukf_seg3.log 2>&1
tail -2 gpurun_out/prof_r02_ukf_seg3.log
