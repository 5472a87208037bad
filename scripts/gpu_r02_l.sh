#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enkf_kernel -s 1 -c 1 -o gpurun_out/prof_r02_enkf_split python scripts/profile_generic.py enkf > gpurun_out/prof_r02_enkf_split.log 2>&1
CDK_ENKF_SPLIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:enkf_kernel -s 1 -c 1 -o gpurun_out/prof_r02_enkf_legacy python scripts/profile_generic.py enkf > gpurun_out/prof_r02_enkf_legacy.log 2>&1
tail -2 gpurun_out/prof_r02_enkf_legacy.log
