#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite except the full-size / long-horizon tests (too slow under the tool)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool ${1:-memcheck} --print-limit 20 python -m pytest tests -q -m gpu -x \
  -k "not full_size and not long_run and not k1000 and not long_runs and not enkf_matches and not default_launch_slices and not c2_kf_n16_k500 and not occupancy_aware and not gradient_long" 2>&1 | tail -8 | tee gpurun_out/r02_${1:-memcheck}_all.txt
