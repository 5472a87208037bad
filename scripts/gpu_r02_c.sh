#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_c_pytest.log
python scripts/strong_probe.py 2>&1 | tee gpurun_out/r02_strong_probe_lean.jsonl
