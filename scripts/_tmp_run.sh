set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r55_pytest.log
timeout 1500 python scripts/bench_configs.py c4 c5 --scale 0.25 2>&1 | tee gpurun_out/r55_configs.jsonl
