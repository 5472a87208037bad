#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python scripts/c3_outputs_probe.py 65536 40000 100000 2>&1 | tee gpurun_out/r02_c3_sliced_probe.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_sliced.json; cut -c1-300 gpurun_out/r02_bench_sliced.json
