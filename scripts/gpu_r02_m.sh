#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02_m_pytest.log
timeout 600 python scripts/bench_configs.py c4 c5 2>&1 | tee gpurun_out/r02_m_c45.jsonl
CDK_GENERIC_REG_ODE=0 timeout 600 python scripts/bench_configs.py c4 2>&1 | tee -a gpurun_out/r02_m_c45.jsonl
