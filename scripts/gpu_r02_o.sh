#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "l96 or L96 or ukf or generic or baseline or golden" 2>&1 | tail -5
timeout 600 python scripts/bench_configs.py c4 2>&1 | tee gpurun_out/r02o_c4.json | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/prof_r02_ukf_rowseg python scripts/profile_generic.py ukf > gpurun_out/prof_r02_ukf_rowseg.log 2>&1
tail -2 gpurun_out/prof_r02_ukf_rowseg.log
