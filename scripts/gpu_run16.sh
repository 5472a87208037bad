set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r16_pytest.log
timeout 1500 python scripts/bench_configs.py c2 c4 c5 --scale 0.25 2>&1 | tee gpurun_out/r16_configs_quarter.jsonl
