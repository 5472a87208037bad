set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/r15_kf python scripts/profile_generic.py kf > gpurun_out/r15_kf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/r15_ukf python scripts/profile_generic.py ukf > gpurun_out/r15_ukf.log 2>&1
tail -3 gpurun_out/r15_kf.log gpurun_out/r15_ukf.log
