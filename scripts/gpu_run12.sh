set -x
timeout 1500 python scripts/bench_configs.py c1 c3 c2 c4 c5 --scale 0.25 2>&1 | tee gpurun_out/r12_configs_quarter.jsonl
