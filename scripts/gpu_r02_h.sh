#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_h_pytest.log
timeout 600 python scripts/bench_configs.py c2 2>&1 | tee gpurun_out/r02_h_c2.jsonl
