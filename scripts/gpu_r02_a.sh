#!/bin/bash
# round 2, call A: micro-benchmark (active lanes), strong-scaling probe, full GPU test suite with error recording
set -x
mkdir -p gpurun_out
./scripts/micro/rk4_pipe > gpurun_out/r02_micro_rk4_lanes.jsonl 2>&1
cat gpurun_out/r02_micro_rk4_lanes.jsonl
python scripts/strong_probe.py 2>&1 | tee gpurun_out/r02_strong_probe_auto.jsonl
CDK_LW_WARPS=1 python scripts/strong_probe.py 8192 16384 2>&1 | tee gpurun_out/r02_strong_probe_w1.jsonl
CDK_LW_WARPS=14 python scripts/strong_probe.py 8192 16384 2>&1 | tee gpurun_out/r02_strong_probe_w14.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r02_a_pytest.log
