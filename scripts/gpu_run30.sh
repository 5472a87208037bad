set -x
mkdir -p gpurun_out
CDK_EKF_MODE=pool timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_small_pool -c 1 -o gpurun_out/r30_pool \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --simple-data > gpurun_out/r30_ncu.log 2>&1
