#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02_k_pytest.log
timeout 600 python scripts/bench_configs.py c5 2>&1 | tee gpurun_out/r02_k_c5.jsonl
CDK_ENKF_CLUSTER=8 timeout 600 python scripts/bench_configs.py c5 2>&1 | tee -a gpurun_out/r02_k_c5.jsonl
CDK_ENKF_SPLIT=0 timeout 600 python scripts/bench_configs.py c5 2>&1 | tee -a gpurun_out/r02_k_c5.jsonl
