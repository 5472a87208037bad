#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/prof_r02_ukf_closed_reg python scripts/profile_generic.py ukf > gpurun_out/prof_r02_ukf_closed_reg.log 2>&1
tail -2 gpurun_out/prof_r02_ukf_closed_reg.log
