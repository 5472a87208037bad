#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "enkf or EnKF or forecast or sample" 2>&1 | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_errors.json'))
for k,v in d.items():
    if 'enkf' in k.lower(): print(k, v)
PY
timeout 600 python scripts/bench_configs.py c5 2>&1 | cut -c1-300
