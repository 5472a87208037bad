#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tests that exercise the kernels changed in round 2
mkdir -p gpurun_out
SEL='time_sliced_launch_is_bit_identical and 3-2-437 or light_mapping_is_bit_identical and 600 or wide_emission or maximum_sizes and 12-33 or test_c4_ukf_l96_n40_m20_vs_oracle and rk4 and False or test_enkf_vs_oracle and l96 or normal_deviates'
for tool in memcheck racecheck; do
  echo "== $tool" | tee -a gpurun_out/r02_sanitizer.txt
  timeout 1700 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -x -q -m gpu -k "$SEL" 2>&1 | tail -25 | tee -a gpurun_out/r02_sanitizer.txt
done
