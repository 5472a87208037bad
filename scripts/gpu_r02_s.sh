#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "l96 or L96 or ekf or forecast or golden" 2>&1 | tail -3
timeout 900 python scripts/c4_solver_probe.py 2048 2>&1 | tee gpurun_out/r02_c4_solver_probe.jsonl
