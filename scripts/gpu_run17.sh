set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r17_pytest.log
timeout 1500 python scripts/bench_configs.py c1 c2 --scale 0.25 2>&1 | tee gpurun_out/r17_configs_quarter.jsonl
CDK_KF_WARP=0 timeout 1500 python scripts/bench_configs.py c2 --scale 0.25 2>&1 | tee gpurun_out/r17_configs_quarter_generic.jsonl
