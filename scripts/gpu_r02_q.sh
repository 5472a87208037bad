#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "enkf or EnKF or sample or forecast or deviates" 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_errors.json'))
for k,v in d.items():
    if 'enkf' in k.lower() or 'sampl' in k.lower() or 'path' in k.lower(): print(k, v)
PY
timeout 600 python scripts/bench_configs.py c5 2>&1 | tee gpurun_out/r02w_c5.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enkf_kernel -c 1 -o gpurun_out/prof_r02_enkf_dg2 python scripts/profile_generic.py enkf > gpurun_out/prof_r02_enkf_dg2.log 2>&1
tail -2 gpurun_out/prof_r02_enkf_dg2.log
