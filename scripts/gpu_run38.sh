set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_warp_filter -s 1 -c 1 -o gpurun_out/r38_kfwarp python scripts/profile_generic.py kf > gpurun_out/r38_kf.log 2>&1
tail -3 gpurun_out/r38_kf.log
