#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tests that exercise the kernels changed in round 2
mkdir -p gpurun_out
rm -f gpurun_out/r02_sanitizer.txt
SEL='time_sliced_launch_is_bit_identical and (437 or 500) or light_mapping_is_bit_identical and 600 or wide_emission or maximum_sizes and 12-33 or register_ode_every_solver and (dopri5 or euler or bosh3) or test_enkf_vs_oracle and l96 or normal_deviates'
for tool in memcheck racecheck; do
  echo "== $tool" | tee -a gpurun_out/r02_sanitizer.txt
  timeout 1700 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -x -q -m gpu -k "$SEL" 2>&1 | tail -8 | tee -a gpurun_out/r02_sanitizer.txt
done
