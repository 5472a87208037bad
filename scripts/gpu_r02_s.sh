#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "l96 or L96 or ukf or forecast or golden" 2>&1 | tail -5
timeout 900 python scripts/c4_solver_probe.py 2048 2>&1 | tee gpurun_out/r02_c4_solver_probe.jsonl
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests -x -q -m gpu -k "register_ode_every_solver and ukf" 2>&1 | tail -4
