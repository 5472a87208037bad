set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/r36_ukf python scripts/profile_generic.py ukf > gpurun_out/r36_ukf.log 2>&1
tail -3 gpurun_out/r36_ukf.log
