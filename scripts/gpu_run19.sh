set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r19_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r19_bench.err | tee gpurun_out/r19_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r19_bench.err | tee gpurun_out/r19_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r19_launches.csv \
  python bench.py --steps 2 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/r19_ncu_bench.log 2>&1
timeout 1500 python scripts/bench_configs.py c1 c2 c4 c5 --scale 0.25 2>&1 | tee gpurun_out/r19_configs_quarter.jsonl
timeout 300 python scripts/trace_lw.py --out gpurun_out/r19_trace_lw.json
timeout 300 python scripts/trace_lw.py --regular --out gpurun_out/r19_trace_lw_regular.json
