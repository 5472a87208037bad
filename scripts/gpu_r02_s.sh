#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python scripts/bench_configs.py c4 c5 2>&1 | cut -c1-330
timeout 900 python scripts/c4_solver_probe.py 2048 2>&1 | head -4
