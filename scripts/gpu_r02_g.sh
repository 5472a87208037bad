#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_g_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/prof_r02_ukf_closed python scripts/profile_generic.py ukf > gpurun_out/prof_r02_ukf_closed.log 2>&1
tail -2 gpurun_out/prof_r02_ukf_closed.log
