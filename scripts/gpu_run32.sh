set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r32_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r32_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r32_bench.err | tee gpurun_out/r32_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r32_bench.err | tee gpurun_out/r32_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r32_launches.csv \
  python bench.py --steps 2 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r32_ncu_bench.log 2>&1
timeout 1500 python scripts/bench_configs.py c1 c2 c3 c4 c5 --scale 0.25 2>&1 | tee gpurun_out/r32_configs_quarter.jsonl
