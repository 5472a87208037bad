set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "kf" 2>&1 | tail -5 | tee gpurun_out/r39_pytest.log
timeout 1500 python scripts/bench_configs.py c2 --scale 0.25 2>&1 | tee gpurun_out/r39_configs.jsonl
timeout 1500 python scripts/bench_configs.py c2 --scale 1.0 2>&1 | tee -a gpurun_out/r39_configs.jsonl
