set -x
CDK_EKF_MODE=warp timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r9_pytest_warp.log
CDK_EKF_MODE=warp timeout 600 python bench.py --steps 5 --warmup 3 --simple-data --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r9_bench_warp.json
CDK_EKF_MODE=lockstep timeout 600 python bench.py --steps 5 --warmup 3 --simple-data --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r9_bench_regroup.json
CDK_EKF_MODE=warp timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_small -s 2 -c 1 -o gpurun_out/r9_ekf_small_warp python bench.py --steps 1 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r9_ncu_full.log 2>&1
