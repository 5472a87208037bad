#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k streaming 2>&1 | tail -8
timeout 900 python scripts/c2_full.py 2>&1 | tail -2 | tee gpurun_out/r02_c2_full.jsonl
timeout 900 python scripts/c2_full.py --solver dopri5 2>&1 | tail -1 | tee -a gpurun_out/r02_c2_full.jsonl
timeout 900 python scripts/c2_full.py --smoother --chunk 8192 2>&1 | tail -1 | tee -a gpurun_out/r02_c2_full.jsonl
free -g | head -2
