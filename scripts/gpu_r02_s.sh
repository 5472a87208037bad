#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/c3_outputs_probe.py 65536 2>&1 | tee -a gpurun_out/r02_c3_outputs_probe.jsonl
timeout 900 python -m pytest tests -x -q -m gpu -k "c3 or l63 or L63 or ekf" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_w8.json; cut -c1-400 gpurun_out/r02_bench_w8.json
