#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02_d_bench.err | tee gpurun_out/r02_d_bench.json
tail -5 gpurun_out/r02_d_bench.err
