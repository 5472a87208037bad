set -x
nvidia-smi -L
python -c "import os; print('cores', os.cpu_count())"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log; cat gpurun_out/r1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r1_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r1_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ekf_small -s 2 -c 1 -o gpurun_out/r1_ekf_small python bench.py --steps 1 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
