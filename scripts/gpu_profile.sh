#!/bin/bash
# One ncu --set full capture of a kernel (read back here with scripts/export_profile.py):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh ekf_small_lw bench'        (the benchmarked kernel)
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh generic_filter ukf'        (scripts/profile_generic.py kf|ukf|enkf)
set -x
kernel=$1; what=${2:-bench}
mkdir -p gpurun_out
if [ "$what" = bench ]; then
  cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --simple-data"; skip=0
else
  cmd="python scripts/profile_generic.py $what"; skip=1
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kernel -s $skip -c 1 -o gpurun_out/prof_${kernel}_${what} $cmd \
  > gpurun_out/prof_${kernel}_${what}.log 2>&1
tail -3 gpurun_out/prof_${kernel}_${what}.log
