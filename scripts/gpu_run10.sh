set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r10_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --simple-data --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r10_bench_warp.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_small -s 2 -c 1 -o gpurun_out/r10_ekf_small_warp python bench.py --steps 1 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r10_ncu_full.log 2>&1
